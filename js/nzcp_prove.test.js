// The proving test the reference lacks (SURVEY.md F1): placed beside /root/reference/test/nzcp.js, it switches the
// snarkjs import for this module and proves + verifies the example pass.  UNTESTED here (no node / circom).
const chai = require("chai");
const path = require("path");
const fs = require("fs");
const snarkjs = require("../js"); // was: require("snarkjs")
const { getToBeSignedAndRs } = require("./helpers/nzcp");
const { bufferToBitArray } = require("./helpers/utils");

const EXAMPLE_PASS_URI = process.env.EXAMPLE_PASS_URI; // the spec pass embedded at test/nzcp.js:51

describe("NZCP Groth16 proof on B200", function () {
  this.timeout(100000);
  it("fullProve + verify (example pass)", async () => {
    const maxLen = 314;
    const data = getToBeSignedAndRs(EXAMPLE_PASS_URI);
    const bytes = Buffer.concat([data.ToBeSigned, Buffer.alloc(maxLen - data.ToBeSigned.length)]);
    const input = { toBeSigned: bufferToBitArray(bytes), toBeSignedLen: data.ToBeSigned.length };
    const wasm = path.join(__dirname, "../circuits/nzcp_exampleTest_js/nzcp_exampleTest.wasm");
    const zkey = path.join(__dirname, "../nzcp_exampleTest_final.zkey");
    const { proof, publicSignals } = await snarkjs.groth16.fullProve(input, wasm, zkey);
    chai.assert.equal(publicSignals.length, 513);
    const vk = await snarkjs.zKey.exportVerificationKey(zkey);
    chai.assert.isTrue(await snarkjs.groth16.verify(vk, publicSignals, proof));
  });
});
