/* libvf_b200_model.so — model-level C-ABI (include/vf_b200_model.h): a thin shim over the CPython interpreter that forwards every call to
 * viewformer_b200/cabi.py with raw pointers.  No arithmetic here; see the header for the rationale. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "vf_b200_model.h"

#ifndef VF_PYTHON_DEFAULT
#define VF_PYTHON_DEFAULT "python3"
#endif
#ifndef VF_REPO_ROOT_DEFAULT
#define VF_REPO_ROOT_DEFAULT ""
#endif

static __thread char g_err[1024];
static PyObject* g_mod = NULL;
static int g_owns_interpreter = 0;

const char* vf_model_last_error(void) { return g_err; }

static void set_err(const char* where) {
    PyObject *t = NULL, *v = NULL, *tb = NULL;
    PyErr_Fetch(&t, &v, &tb);
    PyErr_NormalizeException(&t, &v, &tb);
    const char* msg = "unknown error";
    PyObject* s = v ? PyObject_Str(v) : NULL;
    if (s && PyUnicode_Check(s)) msg = PyUnicode_AsUTF8(s);
    snprintf(g_err, sizeof(g_err), "%s: %s", where, msg ? msg : "unknown error");
    Py_XDECREF(s); Py_XDECREF(t); Py_XDECREF(v); Py_XDECREF(tb);
}

int vf_model_init(void) {
    if (g_mod) return 0;
    if (!Py_IsInitialized()) {
        PyConfig cfg;
        PyConfig_InitPythonConfig(&cfg);
        const char* exe = getenv("VF_PYTHON_EXECUTABLE");
        if (!exe || !exe[0]) exe = VF_PYTHON_DEFAULT;
        PyStatus st = PyConfig_SetBytesString(&cfg, &cfg.executable, exe);       /* a venv interpreter brings its site-packages along */
        if (!PyStatus_Exception(st)) st = Py_InitializeFromConfig(&cfg);
        PyConfig_Clear(&cfg);
        if (PyStatus_Exception(st)) {
            snprintf(g_err, sizeof(g_err), "vf_model_init: cannot start the interpreter (%s)", st.err_msg ? st.err_msg : "?");
            return -1;
        }
        g_owns_interpreter = 1;
    }
    PyGILState_STATE gs = PyGILState_Ensure();
    int rc = 0;
    const char* root = getenv("VF_B200_ROOT");
    if (!root || !root[0]) root = VF_REPO_ROOT_DEFAULT;
    if (root[0]) {                                                               /* make `import viewformer_b200` resolvable */
        PyObject* path = PySys_GetObject("path");
        PyObject* p = PyUnicode_FromString(root);
        if (path && p && !PySequence_Contains(path, p)) PyList_Insert(path, 0, p);
        Py_XDECREF(p);
    }
    g_mod = PyImport_ImportModule("viewformer_b200.cabi");
    if (!g_mod) { set_err("vf_model_init: import viewformer_b200.cabi"); rc = -1; }
    if (g_owns_interpreter && rc == 0) {
        PyGILState_Release(gs);
        PyEval_SaveThread();                                                     /* leave the interpreter unlocked between calls */
        return 0;
    }
    PyGILState_Release(gs);
    return rc;
}

/* call cabi.<fn>(*args) -> long (or tuple of longs into out[]); returns 0 / -1 */
static int call(const char* fn, long long* out, int n_out, const char* fmt, ...) {
    if (vf_model_init() != 0) return -1;
    PyGILState_STATE gs = PyGILState_Ensure();
    int rc = -1;
    va_list ap;
    va_start(ap, fmt);
    PyObject* args = Py_VaBuildValue(fmt, ap);
    va_end(ap);
    PyObject* f = args ? PyObject_GetAttrString(g_mod, fn) : NULL;
    PyObject* r = f ? PyObject_CallObject(f, args) : NULL;
    if (!r) {
        set_err(fn);
    } else {
        rc = 0;
        if (n_out == 1 && PyLong_Check(r)) out[0] = PyLong_AsLongLong(r);
        else if (n_out > 1 && PyTuple_Check(r) && PyTuple_Size(r) >= n_out)
            for (int i = 0; i < n_out; ++i) out[i] = PyLong_AsLongLong(PyTuple_GetItem(r, i));
        else if (n_out > 0) { snprintf(g_err, sizeof(g_err), "%s: unexpected return type", fn); rc = -1; }
    }
    Py_XDECREF(r); Py_XDECREF(f); Py_XDECREF(args);
    PyGILState_Release(gs);
    return rc;
}

#define P(x) ((unsigned long long)(uintptr_t)(x))
#define S(x) ((x) ? (x) : "")

int vf_vq_create(const char* config_json, const char* checkpoint_dir, const char* precision, int device, int64_t seed, vf_handle_t* out) {
    long long h = 0;
    if (!out) { snprintf(g_err, sizeof(g_err), "vf_vq_create: out is NULL"); return -1; }
    int rc = call("vq_create", &h, 1, "(sssiL)", S(config_json), S(checkpoint_dir), S(precision), device, (long long)seed);
    *out = rc == 0 ? (vf_handle_t)h : 0;
    return rc;
}
int vf_vq_info(vf_handle_t h, int* image_size, int* tokens_per_side, int* n_embed, int* in_channels) {
    long long o[4] = {0, 0, 0, 0};
    int rc = call("vq_info", o, 4, "(L)", (long long)h);
    if (rc == 0) {
        if (image_size) *image_size = (int)o[0];
        if (tokens_per_side) *tokens_per_side = (int)o[1];
        if (n_embed) *n_embed = (int)o[2];
        if (in_channels) *in_channels = (int)o[3];
    }
    return rc;
}
int vf_vq_encode(vf_handle_t h, const void* images, int layout, int n, int64_t* codes, vf_cuda_stream_t stream) {
    long long r = 0;
    return call("vq_encode", &r, 1, "(LKiiKK)", (long long)h, P(images), layout, n, P(codes), P(stream));
}
int vf_vq_decode_code(vf_handle_t h, const int64_t* codes, int n, void* images, int layout, vf_cuda_stream_t stream) {
    long long r = 0;
    return call("vq_decode_code", &r, 1, "(LKiKiK)", (long long)h, P(codes), n, P(images), layout, P(stream));
}
int vf_migt_create(const char* config_json, const char* checkpoint_dir, const char* precision, int device, int64_t seed, vf_handle_t* out) {
    long long h = 0;
    if (!out) { snprintf(g_err, sizeof(g_err), "vf_migt_create: out is NULL"); return -1; }
    int rc = call("migt_create", &h, 1, "(sssiL)", S(config_json), S(checkpoint_dir), S(precision), device, (long long)seed);
    *out = rc == 0 ? (vf_handle_t)h : 0;
    return rc;
}
int vf_migt_info(vf_handle_t h, int* tokens_per_side, int* n_embeddings, int* mask_token, int* use_localization) {
    long long o[4] = {0, 0, 0, 0};
    int rc = call("migt_info", o, 4, "(L)", (long long)h);
    if (rc == 0) {
        if (tokens_per_side) *tokens_per_side = (int)o[0];
        if (n_embeddings) *n_embeddings = (int)o[1];
        if (mask_token) *mask_token = (int)o[2];
        if (use_localization) *use_localization = (int)o[3];
    }
    return rc;
}
int vf_migt_forward(vf_handle_t h, const int32_t* input_ids, const float* poses, int B, int T, int64_t* codes_last, float* logits_last,
                    vf_cuda_stream_t stream) {
    long long r = 0;
    return call("migt_forward", &r, 1, "(LKKiiKKK)", (long long)h, P(input_ids), P(poses), B, T, P(codes_last), P(logits_last), P(stream));
}
int vf_migt_prefill_context(vf_handle_t h, const int32_t* context_ids, const float* context_poses, int B, int Tc, vf_handle_t* cache,
                            vf_cuda_stream_t stream) {
    long long c = 0;
    if (!cache) { snprintf(g_err, sizeof(g_err), "vf_migt_prefill_context: cache is NULL"); return -1; }
    int rc = call("migt_prefill_context", &c, 1, "(LKKiiK)", (long long)h, P(context_ids), P(context_poses), B, Tc, P(stream));
    *cache = rc == 0 ? (vf_handle_t)c : 0;
    return rc;
}
int vf_migt_query(vf_handle_t h, vf_handle_t cache, const float* query_poses, int Nq, int64_t* codes, vf_cuda_stream_t stream) {
    long long r = 0;
    return call("migt_query", &r, 1, "(LLKiKK)", (long long)h, (long long)cache, P(query_poses), Nq, P(codes), P(stream));
}
int vf_generate(vf_handle_t transformer, vf_handle_t codebook, const uint8_t* images, const float* cameras, int B, int T,
                uint8_t* generated_images, float* generated_cameras, vf_cuda_stream_t stream) {
    long long r = 0;
    return call("generate", &r, 1, "(LLKKiiKKK)", (long long)transformer, (long long)codebook, P(images), P(cameras), B, T, P(generated_images),
                P(generated_cameras), P(stream));
}
int vf_destroy(vf_handle_t h) {
    long long r = 0;
    return call("destroy", &r, 1, "(L)", (long long)h);
}
