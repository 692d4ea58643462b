// The proving test the reference lacks (SURVEY.md F1): placed beside /root/reference/test/nzcp.js, it builds the input
// exactly as that file does (prepareToBeSigned + bufferToBitArray, test/nzcp.js:11-16, 36-38), switches the snarkjs import
// for this module, proves, checks the 513 public signals the way test/nzcp.js:41-48 checks the witness, and verifies.
// NEVER RUN here (no node / circom / snarkjs in the build environment).
const crypto = require("crypto");
const chai = require("chai");
const path = require("path");
const { verifyPassURIOffline, DID_DOCUMENTS } = require("@vaxxnz/nzcp");
const snarkjs = require("../js"); // was: require("snarkjs")
const { getToBeSignedAndRs } = require("./helpers/nzcp");
const { bufferToBitArray, bitArrayToBuffer } = require("./helpers/utils");

// the spec example pass embedded at /root/reference/test/nzcp.js:51
const EXAMPLE_PASS_URI = "NZCP:/1/2KCEVIQEIVVWK6JNGEASNICZAEP2KALYDZSGSZB2O5SWEOTOPJRXALTDN53GSZBRHEXGQZLBNR2GQLTOPICRUYMBTIFAIGTUKBAAUYTWMOSGQQDDN5XHIZLYOSBHQJTIOR2HA4Z2F4XXO53XFZ3TGLTPOJTS6MRQGE4C6Y3SMVSGK3TUNFQWY4ZPOYYXQKTIOR2HA4Z2F4XW46TDOAXGG33WNFSDCOJONBSWC3DUNAXG46RPMNXW45DFPB2HGL3WGFTXMZLSONUW63TFGEXDALRQMR2HS4DFQJ2FMZLSNFTGSYLCNRSUG4TFMRSW45DJMFWG6UDVMJWGSY2DN53GSZCQMFZXG4LDOJSWIZLOORUWC3CTOVRGUZLDOSRWSZ3JOZSW4TTBNVSWISTBMNVWUZTBNVUWY6KOMFWWKZ2TOBQXE4TPO5RWI33CNIYTSNRQFUYDILJRGYDVAYFE6VGU4MCDGK7DHLLYWHVPUS2YIDJOA6Y524TD3AZRM263WTY2BE4DPKIF27WKF3UDNNVSVWRDYIYVJ65IRJJJ6Z25M2DO4YZLBHWFQGVQR5ZLIWEQJOZTS3IQ7JTNCFDX";

function prepareToBeSigned(input, maxLen) { // test/nzcp.js:11-16
  const bytes = Buffer.alloc(maxLen).fill(0);
  input.copy(bytes, 0);
  return { bytes, bytesLen: input.length };
}

function expectedPubIdentity(passURI) { // test/nzcp.js:18-31
  const res = verifyPassURIOffline(passURI, { didDocument: DID_DOCUMENTS.MOH_EXAMPLE });
  const { givenName, familyName, dob } = res.credentialSubject;
  const toBeSigned = Buffer.from(getToBeSignedAndRs(passURI).ToBeSigned, "hex");
  return {
    credSubjHash: crypto.createHash("sha256").update(`${givenName},${familyName},${dob}`).digest("hex"),
    toBeSignedHash: crypto.createHash("sha256").update(toBeSigned).digest("hex"),
    exp: res.raw.exp,
  };
}

describe("NZCP Groth16 proof on B200", function () {
  this.timeout(100000);
  after(async () => { await snarkjs.terminate(); });

  it("fullProve + verify (example pass)", async () => {
    const maxLen = 314;
    const SHA256_BITS = 256;
    // getToBeSignedAndRs returns a HEX STRING (test/helpers/nzcp.js:154); the reference converts it at test/nzcp.js:36
    const data = prepareToBeSigned(Buffer.from(getToBeSignedAndRs(EXAMPLE_PASS_URI).ToBeSigned, "hex"), maxLen);
    const input = { toBeSigned: bufferToBitArray(data.bytes), toBeSignedLen: data.bytesLen };
    const wasm = path.join(__dirname, "../circuits/nzcp_exampleTest_js/nzcp_exampleTest.wasm");
    const zkey = path.join(__dirname, "../nzcp_exampleTest_final.zkey");
    const { proof, publicSignals } = await snarkjs.groth16.fullProve(input, wasm, zkey);
    chai.assert.equal(publicSignals.length, 2 * SHA256_BITS + 1);
    // the same three checks test/nzcp.js:41-48 makes on witness[1..513]
    const expected = expectedPubIdentity(EXAMPLE_PASS_URI);
    chai.assert.equal(bitArrayToBuffer(publicSignals.slice(0, SHA256_BITS)).toString("hex"), expected.credSubjHash);
    chai.assert.equal(bitArrayToBuffer(publicSignals.slice(SHA256_BITS, 2 * SHA256_BITS)).toString("hex"), expected.toBeSignedHash);
    chai.assert.equal(publicSignals[2 * SHA256_BITS], String(expected.exp));
    const vk = await snarkjs.zKey.exportVerificationKey(zkey);
    chai.assert.isTrue(await snarkjs.groth16.verify(vk, publicSignals, proof));
  });
});
