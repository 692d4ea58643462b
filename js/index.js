// Drop-in for `require("snarkjs")` on the Groth16 proving path (UNTESTED: no node in the build environment).
//
//   const snarkjs = require("nzcp-groth16-b200");            // was: require("snarkjs")
//   const { proof, publicSignals } = await snarkjs.groth16.fullProve(input, wasmFile, zkeyFile);
//
// groth16.prove / fullProve run on the GPU through the N-API addon (nzcp_napi.c -> libnzcp_prover.so);
// witness calculation, verification and everything else are forwarded to the real snarkjs (CPU), which stays a
// dependency.  Same argument conventions as snarkjs 0.4.12: file arguments are a path, a Uint8Array, or
// {type: "mem", data}.  Extension: an optional last argument {r, s, device} injects the blinding scalars
// (32-byte little-endian Buffers) -- snarkjs has no such hook and bit-exact comparison needs one.
"use strict";
const fs = require("fs");
const snarkjs = require("snarkjs");
const addon = require("./nzcp_napi.node");

const cache = new Map(); // zkey identity -> addon handle (proving key resident on the GPU)

function readFileArg(f) {
  if (typeof f === "string") return fs.readFileSync(f);
  if (f && f.type === "mem") return Buffer.from(f.data.buffer, f.data.byteOffset, f.data.byteLength);
  return Buffer.from(f.buffer, f.byteOffset, f.byteLength);
}

function zkeyHandle(zkeyFile, device) {
  let key = zkeyFile;
  if (typeof zkeyFile === "string") {
    const st = fs.statSync(zkeyFile);
    key = `${fs.realpathSync(zkeyFile)}:${st.mtimeMs}:${st.size}:${device}`;
  }
  let h = cache.get(key);
  if (!h) {
    h = addon.zkeyLoad(readFileArg(zkeyFile), device);
    cache.set(key, h);
  }
  return h;
}

function le32ToDec(buf, i) {
  let v = 0n;
  for (let k = 31; k >= 0; k--) v = (v << 8n) | BigInt(buf[32 * i + k]);
  return v.toString();
}

function proofObject(pb) {
  const c = (i) => le32ToDec(pb, i);
  return {
    pi_a: [c(0), c(1), "1"],
    pi_b: [[c(2), c(3)], [c(4), c(5)], ["1", "0"]],
    pi_c: [c(6), c(7), "1"],
    protocol: "groth16",
    curve: "bn128",
  };
}

function publicSignals(wtns, nPublic) {
  // .wtns container: magic, version, nSections, then (u32 id, u64 len, payload)*; section 2 = witness values
  let pos = 12;
  const nsec = wtns.readUInt32LE(8);
  for (let i = 0; i < nsec; i++) {
    const id = wtns.readUInt32LE(pos);
    const len = Number(wtns.readBigUInt64LE(pos + 4));
    pos += 12;
    if (id === 2) {
      const out = [];
      for (let k = 1; k <= nPublic; k++) out.push(le32ToDec(wtns.subarray(pos), k));
      return out;
    }
    pos += len;
  }
  throw new Error("wtns: missing section 2");
}

async function groth16Prove(zkeyFile, witnessFile, logger, opts) {
  opts = opts || {};
  const h = zkeyHandle(zkeyFile, opts.device || 0);
  const wtns = readFileArg(witnessFile);
  if (logger) logger.debug("Proving on GPU");
  const pb = addon.prove(h, wtns, opts.r || null, opts.s || null); // throws with snarkjs's error messages
  return { proof: proofObject(pb), publicSignals: publicSignals(wtns, addon.zkeyInfo(h).nPublic) };
}

async function groth16FullProve(input, wasmFile, zkeyFile, logger, opts) {
  const wtns = { type: "mem" };
  await snarkjs.wtns.calculate(input, wasmFile, wtns); // circom_runtime WitnessCalculator (CPU), as snarkjs does
  return groth16Prove(zkeyFile, wtns, logger, opts);
}

module.exports = Object.assign({}, snarkjs, {
  groth16: Object.assign({}, snarkjs.groth16, { prove: groth16Prove, fullProve: groth16FullProve }),
  terminate: async () => { for (const h of cache.values()) addon.free(h); cache.clear(); },
});
