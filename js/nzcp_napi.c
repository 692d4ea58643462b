/* N-API addon: binds libnzcp_prover.so (include/nzcp_prover.h) for Node.js.
 *
 * STATUS: this build environment has neither node nor node_api.h.  The file is compiled with -Wall -Werror
 * against the hand-declared header js/stub/node_api.h and EXECUTED against the mock N-API runtime js/stub/napi_mock.c
 * (tests/test_napi_shim.py: argument checks and error mapping on CPU, real proofs on the GPU box) -- it has never run
 * under a real node.  js/index.js puts snarkjs's groth16.prove / fullProve surface on top of it.
 *
 * Build (on a box with node >= 12 and the CUDA runtime):
 *   gcc -shared -fPIC -pthread -I$(node -p "require('node:path').dirname(process.execPath)+'/../include/node'") \
 *       -I../include nzcp_napi.c -L../nzcp_circom_b200 -lnzcp_prover -Wl,-rpath,'$ORIGIN/../nzcp_circom_b200' \
 *       -o nzcp_napi.node
 *
 * Exports:
 *   zkeyLoad(Buffer zkey, int device = 0) -> handle                     nzcp_zkey_load + nzcp_prover_create (synchronous:
 *                                                                        a one-time ~1 s expansion of the key on the GPU)
 *   zkeyInfo(handle) -> {nVars, nPublic, domainSize, nCoefs}            nzcp_zkey_info_get
 *   prove(handle, Buffer wtns, Buffer|null r32, Buffer|null s32) -> Promise<Buffer(256)>          nzcp_prove
 *   proveBatch(handle, Buffer[] wtns, Buffer|null r (n*32), Buffer|null s (n*32), int nProvers = 0)
 *                                                         -> Promise<Buffer(n*256)>                nzcp_prove_batch
 *   free(handle)
 * prove / proveBatch run on a libuv worker thread (napi_create_async_work) and settle a Promise, as snarkjs's
 * groth16.prove does -- the event loop is never blocked for the length of a proof.  The handle owns ONE prover
 * (nzcp_prover: one in-flight proof); concurrent prove() calls on the same handle are serialised by a mutex, proveBatch
 * uses the key's own prover pool.  The witness Buffers are referenced for the duration of the work and read in place
 * (pageable memory goes through the library's pinned staging buffer; nothing is copied on the JS side).
 * r / s: null / undefined = random blinding (snarkjs Fr.random()); anything else must be a 32-byte Buffer -- a wrong
 * type or length is a TypeError, never a silent fallback to random.
 * Errors reject the Promise with an Error carrying snarkjs's message (nzcp_last_error(), read on the worker thread).
 */
#include <node_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nzcp_prover.h"

typedef struct {
  nzcp_zkey* zk;
  nzcp_prover* pr;
  pthread_mutex_t mu; /* one in-flight proof per nzcp_prover */
  int in_flight;      /* async works queued or running against this handle */
  int closed;         /* free() was called: released when in_flight drops to zero */
} handle_t;

#define NAPI_OK(call)                                           \
  do {                                                          \
    if ((call) != napi_ok) {                                    \
      napi_throw_error(env, NULL, "N-API call failed: " #call); \
      return NULL;                                              \
    }                                                           \
  } while (0)

static napi_value throw_nzcp(napi_env env) {
  napi_throw_error(env, NULL, nzcp_last_error());
  return NULL;
}

static void handle_release(handle_t* h) {
  if (h->pr) nzcp_prover_free(h->pr);
  if (h->zk) nzcp_zkey_free(h->zk);
  h->pr = NULL;
  h->zk = NULL;
}

static void handle_finalize(napi_env env, void* data, void* hint) {
  handle_t* h = (handle_t*)data;
  (void)env;
  (void)hint;
  if (!h) return;
  handle_release(h);
  pthread_mutex_destroy(&h->mu);
  free(h);
}

static handle_t* get_handle(napi_env env, napi_value v) {
  napi_valuetype t;
  handle_t* h = NULL;
  if (napi_typeof(env, v, &t) != napi_ok || t != napi_external || napi_get_value_external(env, v, (void**)&h) != napi_ok || !h) {
    napi_throw_type_error(env, NULL, "expected a handle returned by zkeyLoad");
    return NULL;
  }
  if (h->closed || !h->zk) {
    napi_throw_error(env, NULL, "the proving key handle was freed");
    return NULL;
  }
  return h;
}

/* Buffer argument -> (data, len); throws a TypeError naming the argument otherwise. */
static int get_buffer(napi_env env, napi_value v, const char* what, void** data, size_t* len) {
  bool is_buf = false;
  if (napi_is_buffer(env, v, &is_buf) != napi_ok || !is_buf || napi_get_buffer_info(env, v, data, len) != napi_ok) {
    char msg[96];
    snprintf(msg, sizeof msg, "%s must be a Buffer", what);
    napi_throw_type_error(env, NULL, msg);
    return 0;
  }
  return 1;
}

/* Optional blinding scalars: null / undefined -> *out = NULL (random); otherwise a Buffer of exactly `bytes` bytes. */
static int get_scalars(napi_env env, napi_value v, const char* what, size_t bytes, const uint8_t** out) {
  napi_valuetype t;
  void* data;
  size_t len;
  *out = NULL;
  if (napi_typeof(env, v, &t) != napi_ok) return 0;
  if (t == napi_null || t == napi_undefined) return 1;
  bool is_buf = false;
  if (napi_is_buffer(env, v, &is_buf) != napi_ok || !is_buf || napi_get_buffer_info(env, v, &data, &len) != napi_ok || len != bytes) {
    char msg[128];
    snprintf(msg, sizeof msg, "%s must be null or a Buffer of %zu bytes (32-byte little-endian scalars)", what, bytes);
    napi_throw_type_error(env, NULL, msg);
    return 0;
  }
  *out = (const uint8_t*)data;
  return 1;
}

static napi_value zkey_load(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  void* data;
  size_t len;
  int32_t device = 0;
  if (argc < 1) {
    napi_throw_type_error(env, NULL, "zkeyLoad(zkey: Buffer, device = 0)");
    return NULL;
  }
  if (!get_buffer(env, argv[0], "zkey", &data, &len)) return NULL;
  if (argc > 1) {
    napi_valuetype t;
    NAPI_OK(napi_typeof(env, argv[1], &t));
    if (t != napi_undefined && napi_get_value_int32(env, argv[1], &device) != napi_ok) {
      napi_throw_type_error(env, NULL, "device must be an integer");
      return NULL;
    }
  }
  handle_t* h = (handle_t*)calloc(1, sizeof *h);
  if (!h) {
    napi_throw_error(env, NULL, "out of memory");
    return NULL;
  }
  pthread_mutex_init(&h->mu, NULL);
  if (nzcp_zkey_load((const uint8_t*)data, len, device, &h->zk) != NZCP_OK || nzcp_prover_create(h->zk, &h->pr) != NZCP_OK) {
    throw_nzcp(env); /* message first: releasing the handle may overwrite the thread's last error */
    handle_finalize(env, h, NULL);
    return NULL;
  }
  napi_value ext;
  if (napi_create_external(env, h, handle_finalize, NULL, &ext) != napi_ok) {
    handle_finalize(env, h, NULL);
    napi_throw_error(env, NULL, "napi_create_external failed");
    return NULL;
  }
  return ext;
}

static napi_value zkey_info(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1], obj, v;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  handle_t* h = argc >= 1 ? get_handle(env, argv[0]) : NULL;
  if (!h) {
    if (argc < 1) napi_throw_type_error(env, NULL, "zkeyInfo(handle)");
    return NULL;
  }
  nzcp_zkey_info zi;
  if (nzcp_zkey_info_get(h->zk, &zi) != NZCP_OK) return throw_nzcp(env);
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_create_uint32(env, zi.n_vars, &v));
  NAPI_OK(napi_set_named_property(env, obj, "nVars", v));
  NAPI_OK(napi_create_uint32(env, zi.n_public, &v));
  NAPI_OK(napi_set_named_property(env, obj, "nPublic", v));
  NAPI_OK(napi_create_uint32(env, zi.domain_size, &v));
  NAPI_OK(napi_set_named_property(env, obj, "domainSize", v));
  NAPI_OK(napi_create_double(env, (double)zi.n_coefs, &v));
  NAPI_OK(napi_set_named_property(env, obj, "nCoefs", v));
  return obj;
}

/* ---- asynchronous proving ------------------------------------------------------------------------------------------ */
typedef struct {
  handle_t* h;
  napi_async_work work;
  napi_deferred deferred;
  size_t n;               /* proofs in this work (1 for prove) */
  int batch, n_provers;
  const uint8_t** wtns;   /* n pointers into the referenced Buffers */
  size_t* wtns_len;
  napi_ref* refs;         /* n + 2 references keeping the witness / r / s Buffers alive */
  size_t n_refs;
  const uint8_t *r, *s;
  nzcp_proof* proofs;     /* n results */
  int rc;
  char err[512];
} job_t;

static void job_free(napi_env env, job_t* j) {
  if (!j) return;
  for (size_t i = 0; i < j->n_refs; i++)
    if (j->refs[i]) napi_delete_reference(env, j->refs[i]);
  if (j->work) napi_delete_async_work(env, j->work);
  free(j->refs);
  free((void*)j->wtns);
  free(j->wtns_len);
  free(j->proofs);
  free(j);
}

static void job_execute(napi_env env, void* data) { /* worker thread: no N-API calls here */
  job_t* j = (job_t*)data;
  (void)env;
  if (j->batch) {
    j->rc = nzcp_prove_batch(j->h->zk, j->wtns, j->wtns_len, j->n, j->r, j->s, j->proofs, j->n_provers, NULL);
  } else {
    pthread_mutex_lock(&j->h->mu);
    j->rc = nzcp_prove(j->h->pr, j->wtns[0], j->wtns_len[0], j->r, j->s, &j->proofs[0], NULL);
    pthread_mutex_unlock(&j->h->mu);
  }
  if (j->rc != NZCP_OK) { /* nzcp_last_error is thread-local: capture it on this thread */
    strncpy(j->err, nzcp_last_error(), sizeof j->err - 1);
    j->err[sizeof j->err - 1] = 0;
  }
}

static void job_complete(napi_env env, napi_status status, void* data) { /* main thread */
  job_t* j = (job_t*)data;
  napi_value val;
  if (status != napi_ok && j->rc == NZCP_OK) {
    j->rc = NZCP_E_INTERNAL;
    snprintf(j->err, sizeof j->err, "proof was cancelled");
  }
  if (j->rc == NZCP_OK) {
    void* dst;
    if (napi_create_buffer_copy(env, j->n * sizeof(nzcp_proof), j->proofs, &dst, &val) == napi_ok)
      napi_resolve_deferred(env, j->deferred, val);
    else
      j->rc = NZCP_E_INTERNAL, snprintf(j->err, sizeof j->err, "napi_create_buffer_copy failed");
  }
  if (j->rc != NZCP_OK) {
    napi_value msg, errv;
    if (napi_create_string_utf8(env, j->err, NAPI_AUTO_LENGTH, &msg) == napi_ok && napi_create_error(env, NULL, msg, &errv) == napi_ok)
      napi_reject_deferred(env, j->deferred, errv);
  }
  handle_t* h = j->h;
  if (--h->in_flight == 0 && h->closed) handle_release(h);
  job_free(env, j);
}

/* Common tail of prove / proveBatch: `wt` holds j->n Buffer values. */
static napi_value job_start(napi_env env, job_t* j, const napi_value* wt, napi_value rv, napi_value sv) {
  napi_value promise, name;
  j->refs = (napi_ref*)calloc(j->n + 2, sizeof(napi_ref));
  j->wtns = (const uint8_t**)calloc(j->n ? j->n : 1, sizeof(uint8_t*));
  j->wtns_len = (size_t*)calloc(j->n ? j->n : 1, sizeof(size_t));
  j->proofs = (nzcp_proof*)calloc(j->n ? j->n : 1, sizeof(nzcp_proof));
  if (!j->refs || !j->wtns || !j->wtns_len || !j->proofs) {
    job_free(env, j);
    napi_throw_error(env, NULL, "out of memory");
    return NULL;
  }
  for (size_t i = 0; i < j->n; i++) {
    void* p;
    if (!get_buffer(env, wt[i], "witness", &p, &j->wtns_len[i])) {
      job_free(env, j);
      return NULL;
    }
    j->wtns[i] = (const uint8_t*)p;
    if (napi_create_reference(env, wt[i], 1, &j->refs[j->n_refs]) != napi_ok) goto fail;
    j->n_refs++;
  }
  if (!get_scalars(env, rv, "r", 32 * j->n, &j->r) || !get_scalars(env, sv, "s", 32 * j->n, &j->s)) {
    job_free(env, j);
    return NULL;
  }
  if (j->r && napi_create_reference(env, rv, 1, &j->refs[j->n_refs]) == napi_ok) j->n_refs++;
  if (j->s && napi_create_reference(env, sv, 1, &j->refs[j->n_refs]) == napi_ok) j->n_refs++;
  if (napi_create_promise(env, &j->deferred, &promise) != napi_ok) goto fail;
  if (napi_create_string_utf8(env, "nzcp_prove", NAPI_AUTO_LENGTH, &name) != napi_ok) goto fail;
  if (napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work) != napi_ok) goto fail;
  j->h->in_flight++;
  if (napi_queue_async_work(env, j->work) != napi_ok) {
    j->h->in_flight--;
    goto fail;
  }
  return promise;
fail:
  job_free(env, j);
  napi_throw_error(env, NULL, "could not start the proving work");
  return NULL;
}

static napi_value prove(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4], undef;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_get_undefined(env, &undef));
  if (argc < 2) {
    napi_throw_type_error(env, NULL, "prove(handle, wtns: Buffer, r?: Buffer(32), s?: Buffer(32))");
    return NULL;
  }
  handle_t* h = get_handle(env, argv[0]);
  if (!h) return NULL;
  job_t* j = (job_t*)calloc(1, sizeof *j);
  if (!j) {
    napi_throw_error(env, NULL, "out of memory");
    return NULL;
  }
  j->h = h;
  j->n = 1;
  return job_start(env, j, &argv[1], argc > 2 ? argv[2] : undef, argc > 3 ? argv[3] : undef);
}

static napi_value prove_batch(napi_env env, napi_callback_info info) {
  size_t argc = 5;
  napi_value argv[5], undef;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_get_undefined(env, &undef));
  if (argc < 2) {
    napi_throw_type_error(env, NULL, "proveBatch(handle, wtns: Buffer[], r?: Buffer(n*32), s?: Buffer(n*32), nProvers = 0)");
    return NULL;
  }
  handle_t* h = get_handle(env, argv[0]);
  if (!h) return NULL;
  bool is_arr = false;
  uint32_t n = 0;
  if (napi_is_array(env, argv[1], &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, argv[1], &n) != napi_ok) {
    napi_throw_type_error(env, NULL, "wtns must be an array of Buffers");
    return NULL;
  }
  int32_t n_provers = 0;
  if (argc > 4) {
    napi_valuetype t;
    NAPI_OK(napi_typeof(env, argv[4], &t));
    if (t != napi_undefined && (napi_get_value_int32(env, argv[4], &n_provers) != napi_ok || n_provers < 0 || n_provers > 64)) {
      napi_throw_type_error(env, NULL, "nProvers must be an integer in [0, 64]");
      return NULL;
    }
  }
  napi_value* wt = (napi_value*)calloc(n ? n : 1, sizeof(napi_value));
  job_t* j = (job_t*)calloc(1, sizeof *j);
  if (!wt || !j) {
    free(wt);
    free(j);
    napi_throw_error(env, NULL, "out of memory");
    return NULL;
  }
  for (uint32_t i = 0; i < n; i++) {
    if (napi_get_element(env, argv[1], i, &wt[i]) != napi_ok) {
      free(wt);
      free(j);
      napi_throw_error(env, NULL, "could not read the witness array");
      return NULL;
    }
  }
  j->h = h;
  j->n = n;
  j->batch = 1;
  j->n_provers = n_provers;
  napi_value res = job_start(env, j, wt, argc > 2 ? argv[2] : undef, argc > 3 ? argv[3] : undef);
  free(wt);
  return res;
}

static napi_value free_handle(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  napi_valuetype t;
  handle_t* h = NULL;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 1 || napi_typeof(env, argv[0], &t) != napi_ok || t != napi_external ||
      napi_get_value_external(env, argv[0], (void**)&h) != napi_ok || !h) {
    napi_throw_type_error(env, NULL, "expected a handle returned by zkeyLoad");
    return NULL;
  }
  h->closed = 1;
  if (h->in_flight == 0) handle_release(h); /* otherwise the last completing work releases it */
  return NULL;
}

static napi_value init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"zkeyLoad", NULL, zkey_load, NULL, NULL, NULL, napi_default, NULL},
      {"zkeyInfo", NULL, zkey_info, NULL, NULL, NULL, napi_default, NULL},
      {"prove", NULL, prove, NULL, NULL, NULL, napi_default, NULL},
      {"proveBatch", NULL, prove_batch, NULL, NULL, NULL, napi_default, NULL},
      {"free", NULL, free_handle, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, init)
