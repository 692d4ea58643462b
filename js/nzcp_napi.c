/* N-API addon: binds libnzcp_prover.so (include/nzcp_prover.h) 1:1 for Node.js.
 *
 * STATUS: written against the N-API C interface from memory; this build environment has neither node nor
 * node_api.h, so this file has NEVER been compiled or run.  It shows the reference-side binding a maintainer adds
 * (see INTEGRATION.md); the tested host mirror in this repo is the Python ctypes one (nzcp_circom_b200/).
 *
 * Build (on a box with node >= 12 and the CUDA runtime):
 *   gcc -shared -fPIC -I$(node -p "require('node:path').dirname(process.execPath)+'/../include/node'") \
 *       -I../include nzcp_napi.c -L../nzcp_circom_b200 -lnzcp_prover -Wl,-rpath,'$ORIGIN/../nzcp_circom_b200' \
 *       -o nzcp_napi.node
 *
 * Exports:
 *   zkeyLoad(Buffer zkey, int device) -> external handle          nzcp_zkey_load + nzcp_prover_create
 *   zkeyInfo(handle) -> {nVars, nPublic, domainSize}
 *   prove(handle, Buffer wtns, Buffer|null r32, Buffer|null s32) -> Buffer(256)   nzcp_prove
 *   free(handle)
 * Buffers are passed by pointer -- the library reads the .zkey / .wtns sections straight out of the (pinned, when
 * the caller allocated them with cudaHostRegister) node Buffers; nothing is copied on the JS side.
 * Errors become JS exceptions carrying snarkjs's messages (nzcp_last_error()).
 */
#include <node_api.h>
#include <stdlib.h>
#include <string.h>

#include "nzcp_prover.h"

typedef struct {
  nzcp_zkey* zk;
  nzcp_prover* pr;
} handle_t;

#define NAPI_OK(call)                                          \
  do {                                                         \
    if ((call) != napi_ok) {                                   \
      napi_throw_error(env, NULL, "N-API call failed: " #call); \
      return NULL;                                             \
    }                                                          \
  } while (0)

static napi_value throw_nzcp(napi_env env) {
  napi_throw_error(env, NULL, nzcp_last_error());
  return NULL;
}

static void handle_finalize(napi_env env, void* data, void* hint) {
  handle_t* h = (handle_t*)data;
  (void)env;
  (void)hint;
  if (!h) return;
  if (h->pr) nzcp_prover_free(h->pr);
  if (h->zk) nzcp_zkey_free(h->zk);
  free(h);
}

static napi_value zkey_load(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  void* data;
  size_t len;
  int32_t device = 0;
  NAPI_OK(napi_get_buffer_info(env, argv[0], &data, &len));
  if (argc > 1) NAPI_OK(napi_get_value_int32(env, argv[1], &device));
  handle_t* h = (handle_t*)calloc(1, sizeof *h);
  if (nzcp_zkey_load((const uint8_t*)data, len, device, &h->zk) != NZCP_OK ||
      nzcp_prover_create(h->zk, &h->pr) != NZCP_OK) {
    handle_finalize(env, h, NULL);
    return throw_nzcp(env);
  }
  napi_value ext;
  NAPI_OK(napi_create_external(env, h, handle_finalize, NULL, &ext));
  return ext;
}

static napi_value zkey_info(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1], obj, v;
  handle_t* h;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_get_value_external(env, argv[0], (void**)&h));
  nzcp_zkey_info zi;
  if (nzcp_zkey_info_get(h->zk, &zi) != NZCP_OK) return throw_nzcp(env);
  NAPI_OK(napi_create_object(env, &obj));
  NAPI_OK(napi_create_uint32(env, zi.n_vars, &v));
  NAPI_OK(napi_set_named_property(env, obj, "nVars", v));
  NAPI_OK(napi_create_uint32(env, zi.n_public, &v));
  NAPI_OK(napi_set_named_property(env, obj, "nPublic", v));
  NAPI_OK(napi_create_uint32(env, zi.domain_size, &v));
  NAPI_OK(napi_set_named_property(env, obj, "domainSize", v));
  return obj;
}

static const uint8_t* opt_scalar(napi_env env, napi_value v) {
  napi_valuetype t;
  void* data;
  size_t len;
  if (napi_typeof(env, v, &t) != napi_ok || t == napi_null || t == napi_undefined) return NULL;
  if (napi_get_buffer_info(env, v, &data, &len) != napi_ok || len != 32) return NULL;
  return (const uint8_t*)data;
}

static napi_value prove(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4], out;
  handle_t* h;
  void* wt;
  size_t wlen;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_get_value_external(env, argv[0], (void**)&h));
  NAPI_OK(napi_get_buffer_info(env, argv[1], &wt, &wlen));
  const uint8_t* r = argc > 2 ? opt_scalar(env, argv[2]) : NULL;
  const uint8_t* s = argc > 3 ? opt_scalar(env, argv[3]) : NULL;
  nzcp_proof pf;
  if (nzcp_prove(h->pr, (const uint8_t*)wt, wlen, r, s, &pf, NULL) != NZCP_OK) return throw_nzcp(env);
  void* dst;
  NAPI_OK(napi_create_buffer_copy(env, sizeof pf, &pf, &dst, &out));
  return out;
}

static napi_value free_handle(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  handle_t* h;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  NAPI_OK(napi_get_value_external(env, argv[0], (void**)&h));
  if (h->pr) nzcp_prover_free(h->pr);
  if (h->zk) nzcp_zkey_free(h->zk);
  h->pr = NULL;
  h->zk = NULL;
  return NULL;
}

static napi_value init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"zkeyLoad", NULL, zkey_load, NULL, NULL, NULL, napi_default, NULL},
      {"zkeyInfo", NULL, zkey_info, NULL, NULL, NULL, napi_default, NULL},
      {"prove", NULL, prove, NULL, NULL, NULL, napi_default, NULL},
      {"free", NULL, free_handle, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, init)
