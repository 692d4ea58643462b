/* STUB of Node's <node_api.h>, hand-declared from the documented N-API (Node-API version 6) C interface.
 *
 * This build environment has neither node nor its headers.  The stub exists so that js/nzcp_napi.c is
 * at least COMPILED (tests/test_napi_shim.py: gcc -Wall -Werror against this header) and EXECUTED against the mock
 * runtime in napi_mock.c -- both catch shim bugs (argument handling, async flow, error mapping, buffer lifetimes) long
 * before a box with node is available.  It declares only what the addon uses.  On a real build the compiler's include
 * path puts Node's own header first and this file is never seen.  NOT a substitute for testing under node.
 */
#ifndef NZCP_STUB_NODE_API_H
#define NZCP_STUB_NODE_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAPI_AUTO_LENGTH SIZE_MAX

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_deferred__* napi_deferred;
typedef struct napi_async_work__* napi_async_work;
typedef struct napi_callback_info__* napi_callback_info;

typedef enum {
  napi_ok,
  napi_invalid_arg,
  napi_object_expected,
  napi_string_expected,
  napi_name_expected,
  napi_function_expected,
  napi_number_expected,
  napi_boolean_expected,
  napi_array_expected,
  napi_generic_failure,
  napi_pending_exception,
  napi_cancelled
} napi_status;

typedef enum {
  napi_undefined,
  napi_null,
  napi_boolean,
  napi_number,
  napi_string,
  napi_symbol,
  napi_object,
  napi_function,
  napi_external,
  napi_bigint
} napi_valuetype;

typedef enum { napi_default = 0, napi_writable = 1 << 0, napi_enumerable = 1 << 1, napi_configurable = 1 << 2 } napi_property_attributes;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void* data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void* data);

typedef struct {
  const char* utf8name;
  napi_value name;
  napi_callback method;
  napi_callback getter;
  napi_callback setter;
  napi_value value;
  napi_property_attributes attributes;
  void* data;
} napi_property_descriptor;

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);

typedef struct napi_module {
  int nm_version;
  unsigned int nm_flags;
  const char* nm_filename;
  napi_addon_register_func nm_register_func;
  const char* nm_modname;
  void* nm_priv;
  void* reserved[4];
} napi_module;

void napi_module_register(napi_module* mod);

#define NAPI_MODULE_X(modname, regfunc, priv, flags)                                              \
  static napi_module _module = {1, flags, __FILE__, regfunc, #modname, priv, {0, 0, 0, 0}};       \
  static void _register_##modname(void) __attribute__((constructor));                              \
  static void _register_##modname(void) { napi_module_register(&_module); }
#define NAPI_MODULE(modname, regfunc) NAPI_MODULE_X(modname, regfunc, NULL, 0)

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_get_undefined(napi_env env, napi_value* result);
napi_status napi_get_null(napi_env env, napi_value* result);
napi_status napi_is_buffer(napi_env env, napi_value value, bool* result);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_create_buffer_copy(napi_env env, size_t length, const void* data, void** result_data, napi_value* result);
napi_status napi_is_array(napi_env env, napi_value value, bool* result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_create_uint32(napi_env env, uint32_t value, napi_value* result);
napi_status napi_create_double(napi_env env, double value, napi_value* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env env, napi_ref ref);
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_create_promise(napi_env env, napi_deferred* deferred, napi_value* promise);
napi_status napi_resolve_deferred(napi_env env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete, void* data,
                                   napi_async_work* result);
napi_status napi_queue_async_work(napi_env env, napi_async_work work);
napi_status napi_delete_async_work(napi_env env, napi_async_work work);

#ifdef __cplusplus
}
#endif
#endif
