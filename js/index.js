// Drop-in for `require("snarkjs")` on the Groth16 proving path.  NEVER RUN: there is no node in the build environment;
// the addon underneath (nzcp_napi.c) is compiled and driven through a mock N-API runtime (tests/test_napi_shim.py), this
// file is not.
//
//   const snarkjs = require("nzcp-groth16-b200");            // was: require("snarkjs")
//   const { proof, publicSignals } = await snarkjs.groth16.fullProve(input, wasmFile, zkeyFile);
//
// groth16.prove / fullProve run on the GPU through the N-API addon (nzcp_napi.c -> libnzcp_prover.so);
// witness calculation, verification and everything else are forwarded to the real snarkjs (CPU), which stays a
// dependency.  Same argument conventions as snarkjs 0.4.12: file arguments are a path, a Uint8Array, or
// {type: "mem", data}.  Extensions: an optional last argument {r, s, device} injects the blinding scalars (BigInt,
// decimal string or 32-byte little-endian Buffer) -- snarkjs has no such hook and bit-exact comparison needs one;
// groth16.proveBatch(zkey, [wtns...]) keeps several proofs in flight on the GPU (nzcp_prove_batch).
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

// Blinding scalar -> 32-byte little-endian Buffer; null / undefined -> null (random, as snarkjs).  Anything else throws:
// a caller that injects r, s for a bit-exact comparison must never get a silently random proof.
function scalar32(v, name) {
  if (v === null || v === undefined) return null;
  if (Buffer.isBuffer(v) || v instanceof Uint8Array) {
    if (v.length !== 32) throw new TypeError(`${name} must be 32 bytes`);
    return Buffer.from(v.buffer, v.byteOffset, 32);
  }
  if (typeof v === "bigint" || typeof v === "string" || typeof v === "number") {
    let x = BigInt(v);
    if (x < 0n || x >> 256n) throw new RangeError(`${name} out of range`);
    const out = Buffer.alloc(32);
    for (let k = 0; k < 32; k++, x >>= 8n) out[k] = Number(x & 0xffn);
    return out;
  }
  throw new TypeError(`${name} must be a BigInt, a decimal string or a 32-byte Buffer`);
}

function proofObject(pb, off = 0) {
  const c = (i) => le32ToDec(pb, off / 32 + i);
  const zero = (i, n) => pb.subarray(off + 32 * i, off + 32 * (i + n)).every((b) => b === 0);
  // the point at infinity prints as ffjavascript's G.zero: [0, 1, 0]
  return {
    pi_a: zero(0, 2) ? ["0", "1", "0"] : [c(0), c(1), "1"],
    pi_b: zero(2, 4) ? [["0", "0"], ["1", "0"], ["0", "0"]] : [[c(2), c(3)], [c(4), c(5)], ["1", "0"]],
    pi_c: zero(6, 2) ? ["0", "1", "0"] : [c(6), c(7), "1"],
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
  // the addon proves on a worker thread and rejects with snarkjs's own error messages
  const pb = await addon.prove(h, wtns, scalar32(opts.r, "r"), scalar32(opts.s, "s"));
  return { proof: proofObject(pb), publicSignals: publicSignals(wtns, addon.zkeyInfo(h).nPublic) };
}

async function groth16FullProve(input, wasmFile, zkeyFile, logger, opts) {
  const wtns = { type: "mem" };
  await snarkjs.wtns.calculate(input, wasmFile, wtns); // circom_runtime WitnessCalculator (CPU), as snarkjs does
  return groth16Prove(zkeyFile, wtns, logger, opts);
}

// Not in snarkjs: many witnesses against one key, several proofs in flight on the GPU.  opts.r / opts.s: arrays.
async function groth16ProveBatch(zkeyFile, witnessFiles, logger, opts) {
  opts = opts || {};
  const h = zkeyHandle(zkeyFile, opts.device || 0);
  const wt = witnessFiles.map(readFileArg);
  const cat = (a, name) => (a ? Buffer.concat(a.map((v, i) => scalar32(v, `${name}[${i}]`))) : null);
  const pb = await addon.proveBatch(h, wt, cat(opts.r, "r"), cat(opts.s, "s"), opts.provers || 0);
  const nPublic = addon.zkeyInfo(h).nPublic;
  return wt.map((w, i) => ({ proof: proofObject(pb, 256 * i), publicSignals: publicSignals(w, nPublic) }));
}

module.exports = Object.assign({}, snarkjs, {
  groth16: Object.assign({}, snarkjs.groth16, { prove: groth16Prove, fullProve: groth16FullProve, proveBatch: groth16ProveBatch }),
  terminate: async () => { for (const h of cache.values()) addon.free(h); cache.clear(); },
});
