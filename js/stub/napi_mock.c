/* Mock N-API runtime: just enough of Node's C interface (js/stub/node_api.h) to LOAD and DRIVE js/nzcp_napi.c without
 * node -- test infrastructure, not product code.  tests/test_napi_shim.py links nzcp_napi.c + this file + libnzcp_prover.so
 * into one shared object and calls the addon's exported functions through the mock_* driver functions below (ctypes).
 *
 * Fidelity notes: values are heap objects that live until mock_reset(); a thrown exception is recorded and the next
 * mock_call returns NULL (like a JS call that threw); napi_queue_async_work runs `execute` on a fresh pthread (so that
 * thread-local state such as nzcp_last_error() behaves as under libuv) and `complete` on the calling thread when
 * mock_run_pending() is called -- the driver's stand-in for returning to the event loop.
 */
#include <node_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef enum { V_UNDEF, V_NULL, V_NUMBER, V_STRING, V_BUFFER, V_EXTERNAL, V_OBJECT, V_ARRAY, V_FUNCTION, V_PROMISE, V_ERROR } vkind;

typedef struct prop {
  char* name;
  struct napi_value__* val;
  struct prop* next;
} prop;

struct napi_value__ {
  vkind kind;
  double num;
  char* str;            /* V_STRING, V_ERROR (message) */
  void* ptr;            /* V_BUFFER data, V_EXTERNAL data */
  size_t len;
  int owned;            /* V_BUFFER: free(ptr) at reset */
  napi_finalize fin;    /* V_EXTERNAL */
  void* fin_hint;
  prop* props;          /* V_OBJECT */
  struct napi_value__** elems; /* V_ARRAY */
  napi_callback fn;     /* V_FUNCTION */
  int state;            /* V_PROMISE: 0 pending, 1 resolved, 2 rejected */
  struct napi_value__* settled;
  int is_type_error;    /* V_ERROR */
  struct napi_value__* next_all;
};

struct napi_env__ { int unused; };
struct napi_ref__ { napi_value v; };
struct napi_deferred__ { napi_value promise; };
struct napi_callback_info__ { size_t argc; napi_value* argv; };
struct napi_async_work__ {
  napi_async_execute_callback execute;
  napi_async_complete_callback complete;
  void* data;
  pthread_t thread;
  int running;
  struct napi_async_work__* next;
};

static struct napi_env__ g_env;
static napi_value g_all = NULL;
static napi_value g_exception = NULL;
static napi_module* g_module = NULL;
static struct napi_async_work__* g_pending = NULL;
static int g_live_refs = 0, g_live_works = 0;

static napi_value mk(vkind k) {
  napi_value v = (napi_value)calloc(1, sizeof *v);
  v->kind = k;
  v->next_all = g_all;
  g_all = v;
  return v;
}

void napi_module_register(napi_module* mod) { g_module = mod; }

napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
  (void)env;
  if (argc) {
    size_t cap = *argc;
    for (size_t i = 0; i < cap; i++) argv[i] = i < info->argc ? info->argv[i] : mk(V_UNDEF);
    *argc = info->argc;
  }
  if (this_arg) *this_arg = mk(V_UNDEF);
  if (data) *data = NULL;
  return napi_ok;
}

napi_status napi_typeof(napi_env env, napi_value v, napi_valuetype* r) {
  (void)env;
  if (!v) return napi_invalid_arg;
  switch (v->kind) {
    case V_UNDEF: *r = napi_undefined; break;
    case V_NULL: *r = napi_null; break;
    case V_NUMBER: *r = napi_number; break;
    case V_STRING: *r = napi_string; break;
    case V_EXTERNAL: *r = napi_external; break;
    case V_FUNCTION: *r = napi_function; break;
    default: *r = napi_object; break;   /* Buffers, arrays, promises, errors are objects */
  }
  return napi_ok;
}

napi_status napi_get_undefined(napi_env env, napi_value* r) { (void)env; *r = mk(V_UNDEF); return napi_ok; }
napi_status napi_get_null(napi_env env, napi_value* r) { (void)env; *r = mk(V_NULL); return napi_ok; }
napi_status napi_is_buffer(napi_env env, napi_value v, bool* r) { (void)env; *r = v && v->kind == V_BUFFER; return napi_ok; }
napi_status napi_get_buffer_info(napi_env env, napi_value v, void** data, size_t* len) {
  (void)env;
  if (!v || v->kind != V_BUFFER) return napi_invalid_arg;
  if (data) *data = v->ptr;
  if (len) *len = v->len;
  return napi_ok;
}
napi_status napi_create_buffer_copy(napi_env env, size_t length, const void* data, void** result_data, napi_value* result) {
  (void)env;
  napi_value v = mk(V_BUFFER);
  v->ptr = malloc(length ? length : 1);
  memcpy(v->ptr, data, length);
  v->len = length;
  v->owned = 1;
  if (result_data) *result_data = v->ptr;
  *result = v;
  return napi_ok;
}
napi_status napi_is_array(napi_env env, napi_value v, bool* r) { (void)env; *r = v && v->kind == V_ARRAY; return napi_ok; }
napi_status napi_get_array_length(napi_env env, napi_value v, uint32_t* r) {
  (void)env;
  if (!v || v->kind != V_ARRAY) return napi_array_expected;
  *r = (uint32_t)v->len;
  return napi_ok;
}
napi_status napi_get_element(napi_env env, napi_value v, uint32_t i, napi_value* r) {
  (void)env;
  if (!v || v->kind != V_ARRAY) return napi_array_expected;
  *r = i < v->len ? v->elems[i] : mk(V_UNDEF);
  return napi_ok;
}
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t* r) {
  (void)env;
  if (!v || v->kind != V_NUMBER) return napi_number_expected;
  *r = (int32_t)v->num;
  return napi_ok;
}
napi_status napi_get_value_uint32(napi_env env, napi_value v, uint32_t* r) {
  (void)env;
  if (!v || v->kind != V_NUMBER) return napi_number_expected;
  *r = (uint32_t)v->num;
  return napi_ok;
}
napi_status napi_create_uint32(napi_env env, uint32_t x, napi_value* r) { (void)env; *r = mk(V_NUMBER); (*r)->num = x; return napi_ok; }
napi_status napi_create_double(napi_env env, double x, napi_value* r) { (void)env; *r = mk(V_NUMBER); (*r)->num = x; return napi_ok; }
napi_status napi_create_string_utf8(napi_env env, const char* s, size_t len, napi_value* r) {
  (void)env;
  if (len == NAPI_AUTO_LENGTH) len = strlen(s);
  *r = mk(V_STRING);
  (*r)->str = (char*)malloc(len + 1);
  memcpy((*r)->str, s, len);
  (*r)->str[len] = 0;
  return napi_ok;
}
napi_status napi_create_object(napi_env env, napi_value* r) { (void)env; *r = mk(V_OBJECT); return napi_ok; }
napi_status napi_set_named_property(napi_env env, napi_value o, const char* name, napi_value val) {
  (void)env;
  if (!o || o->kind != V_OBJECT) return napi_object_expected;
  prop* p = (prop*)calloc(1, sizeof *p);
  p->name = strdup(name);
  p->val = val;
  p->next = o->props;
  o->props = p;
  return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value o, size_t n, const napi_property_descriptor* d) {
  for (size_t i = 0; i < n; i++) {
    napi_value f = d[i].value;
    if (d[i].method) {
      f = mk(V_FUNCTION);
      f->fn = d[i].method;
    }
    napi_set_named_property(env, o, d[i].utf8name, f);
  }
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void* hint, napi_value* r) {
  (void)env;
  *r = mk(V_EXTERNAL);
  (*r)->ptr = data;
  (*r)->fin = fin;
  (*r)->fin_hint = hint;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env env, napi_value v, void** r) {
  (void)env;
  if (!v || v->kind != V_EXTERNAL) return napi_invalid_arg;
  *r = v->ptr;
  return napi_ok;
}
napi_status napi_create_reference(napi_env env, napi_value v, uint32_t rc, napi_ref* r) {
  (void)env;
  (void)rc;
  *r = (napi_ref)calloc(1, sizeof **r);
  (*r)->v = v;
  g_live_refs++;
  return napi_ok;
}
napi_status napi_delete_reference(napi_env env, napi_ref r) { (void)env; free(r); g_live_refs--; return napi_ok; }
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* r) {
  (void)env;
  (void)code;
  if (!msg || msg->kind != V_STRING) return napi_string_expected;
  *r = mk(V_ERROR);
  (*r)->str = strdup(msg->str);
  return napi_ok;
}
static napi_status throw_kind(const char* msg, int type_error) {
  if (g_exception) return napi_pending_exception;
  g_exception = mk(V_ERROR);
  g_exception->str = strdup(msg ? msg : "");
  g_exception->is_type_error = type_error;
  return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) { (void)env; (void)code; return throw_kind(msg, 0); }
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg) { (void)env; (void)code; return throw_kind(msg, 1); }
napi_status napi_create_promise(napi_env env, napi_deferred* d, napi_value* promise) {
  (void)env;
  *promise = mk(V_PROMISE);
  *d = (napi_deferred)calloc(1, sizeof **d);
  (*d)->promise = *promise;
  return napi_ok;
}
static napi_status settle(napi_deferred d, napi_value v, int state) {
  if (!d || d->promise->state) return napi_generic_failure;
  d->promise->state = state;
  d->promise->settled = v;
  free(d);
  return napi_ok;
}
napi_status napi_resolve_deferred(napi_env env, napi_deferred d, napi_value v) { (void)env; return settle(d, v, 1); }
napi_status napi_reject_deferred(napi_env env, napi_deferred d, napi_value v) { (void)env; return settle(d, v, 2); }

napi_status napi_create_async_work(napi_env env, napi_value res, napi_value name, napi_async_execute_callback ex,
                                   napi_async_complete_callback done, void* data, napi_async_work* r) {
  (void)env;
  (void)res;
  (void)name;
  *r = (napi_async_work)calloc(1, sizeof **r);
  (*r)->execute = ex;
  (*r)->complete = done;
  (*r)->data = data;
  g_live_works++;
  return napi_ok;
}
static void* work_thread(void* p) {
  napi_async_work w = (napi_async_work)p;
  w->execute(&g_env, w->data);
  return NULL;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) {
  (void)env;
  if (pthread_create(&w->thread, NULL, work_thread, w) != 0) return napi_generic_failure;
  w->running = 1;
  w->next = g_pending;
  g_pending = w;
  return napi_ok;
}
napi_status napi_delete_async_work(napi_env env, napi_async_work w) { (void)env; free(w); g_live_works--; return napi_ok; }

/* ------------------------------------------------------------------------------------------------- driver (ctypes) */
napi_value mock_init(void) { /* what `require("./nzcp_napi.node")` does: run the registered init on a fresh exports object */
  if (!g_module) return NULL;
  napi_value exports = mk(V_OBJECT);
  return g_module->nm_register_func(&g_env, exports);
}
napi_value mock_undefined(void) { return mk(V_UNDEF); }
napi_value mock_null(void) { return mk(V_NULL); }
napi_value mock_number(double x) { napi_value v = mk(V_NUMBER); v->num = x; return v; }
napi_value mock_string(const char* s) { napi_value v; napi_create_string_utf8(&g_env, s, NAPI_AUTO_LENGTH, &v); return v; }
napi_value mock_buffer(void* data, size_t len) { /* a Buffer over caller-owned memory (kept alive by the test) */
  napi_value v = mk(V_BUFFER);
  v->ptr = data;
  v->len = len;
  return v;
}
napi_value mock_array(napi_value* elems, size_t n) {
  napi_value v = mk(V_ARRAY);
  v->elems = (napi_value*)calloc(n ? n : 1, sizeof(napi_value));
  memcpy(v->elems, elems, n * sizeof(napi_value));
  v->len = n;
  return v;
}
napi_value mock_get(napi_value obj, const char* name) {
  if (!obj || obj->kind != V_OBJECT) return NULL;
  for (prop* p = obj->props; p; p = p->next)
    if (strcmp(p->name, name) == 0) return p->val;
  return NULL;
}
/* Call exports[name](argv...).  Returns NULL if the function threw; mock_exception_* then describe the exception. */
napi_value mock_call(napi_value exports, const char* name, napi_value* argv, size_t argc) {
  napi_value f = mock_get(exports, name);
  if (!f || f->kind != V_FUNCTION) return NULL;
  g_exception = NULL;
  struct napi_callback_info__ info = {argc, argv};
  napi_value r = f->fn(&g_env, &info);
  if (g_exception) return NULL;
  return r ? r : mk(V_UNDEF);
}
const char* mock_exception_message(void) { return g_exception ? g_exception->str : NULL; }
int mock_exception_is_type_error(void) { return g_exception ? g_exception->is_type_error : 0; }
/* Back to the "event loop": join every queued work's thread and run its complete callback on this thread. */
int mock_run_pending(void) {
  int n = 0;
  while (g_pending) {
    napi_async_work w = g_pending;
    g_pending = w->next;
    pthread_join(w->thread, NULL);
    w->complete(&g_env, napi_ok, w->data);
    n++;
  }
  return n;
}
int mock_kind(napi_value v) { return v ? (int)v->kind : -1; }
double mock_number_value(napi_value v) { return v && v->kind == V_NUMBER ? v->num : 0; }
int mock_promise_state(napi_value v) { return v && v->kind == V_PROMISE ? v->state : -1; }
napi_value mock_promise_value(napi_value v) { return v && v->kind == V_PROMISE ? v->settled : NULL; }
const char* mock_error_message(napi_value v) { return v && v->kind == V_ERROR ? v->str : NULL; }
size_t mock_buffer_len(napi_value v) { return v && v->kind == V_BUFFER ? v->len : 0; }
void* mock_buffer_data(napi_value v) { return v && v->kind == V_BUFFER ? v->ptr : NULL; }
int mock_live_refs(void) { return g_live_refs; }
int mock_live_works(void) { return g_live_works; }
/* Garbage-collect everything: externals are finalized (the addon frees its GPU state there), all values are freed. */
void mock_reset(void) {
  mock_run_pending();
  for (napi_value v = g_all; v;) {
    napi_value nx = v->next_all;
    if (v->kind == V_EXTERNAL && v->fin) v->fin(&g_env, v->ptr, v->fin_hint);
    if (v->kind == V_BUFFER && v->owned) free(v->ptr);
    for (prop* p = v->props; p;) {
      prop* pn = p->next;
      free(p->name);
      free(p);
      p = pn;
    }
    free(v->elems);
    free(v->str);
    free(v);
    v = nx;
  }
  g_all = NULL;
  g_exception = NULL;
}
