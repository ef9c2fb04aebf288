/* MINIMAL N-API DECLARATIONS FOR TYPE CHECKING ONLY.
 * This image has neither node nor its headers; this file declares just the Node-API (stable C ABI, nodejs.org/api/n-api)
 * types and functions addon/binding.cc uses -- with the parameter lists of Node's own js_native_api.h / node_api.h
 * (Node-API version 8) -- so that `g++ -fsyntax-only` can check the shim here. It proves the C++ type-checks against
 * those signatures, not that it loads: a real build uses node's own <node_api.h> (node-gyp / cmake-js put it on the
 * include path) and never sees this file. */
#ifndef GVT_NODE_API_STUB_H
#define GVT_NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_ref__* napi_ref;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_generic_failure = 9 } napi_status;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef enum {
    napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
    napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
typedef struct {
    const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
    napi_value value; napi_property_attributes attributes; void* data;
} napi_property_descriptor;
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_wrap(napi_env, napi_value js_object, void* native_object, napi_finalize, void* hint, napi_ref* result);
napi_status napi_unwrap(napi_env, napi_value js_object, void** result);
napi_status napi_get_value_double(napi_env, napi_value, double* result);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t* result);
napi_status napi_get_value_bool(napi_env, napi_value, bool* result);
napi_status napi_create_double(napi_env, double value, napi_value* result);
napi_status napi_create_uint32(napi_env, uint32_t value, napi_value* result);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_get_undefined(napi_env, napi_value* result);
napi_status napi_get_element(napi_env, napi_value object, uint32_t index, napi_value* result);
napi_status napi_get_array_length(napi_env, napi_value, uint32_t* result);
napi_status napi_get_named_property(napi_env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* utf8name, bool* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_get_arraybuffer_info(napi_env, napi_value arraybuffer, void** data, size_t* byte_length);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value* result);
napi_status napi_get_typedarray_info(napi_env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_define_class(napi_env, const char* utf8name, size_t length, napi_callback constructor, void* data, size_t property_count, const napi_property_descriptor* properties, napi_value* result);
napi_status napi_is_typedarray(napi_env, napi_value value, bool* result);
napi_status napi_is_arraybuffer(napi_env, napi_value value, bool* result);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref* result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
napi_status napi_get_reference_value(napi_env, napi_ref ref, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_throw_range_error(napi_env, const char* code, const char* msg);
#define NAPI_MODULE_INIT() extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports)

/* externals (opaque native pointers held by JS values) */
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
#ifdef __cplusplus
}
#endif
#endif
