// `run` and `train` of libpybindings.so (/root/reference/src/pybindings.h:16-27, src/pybindings.cpp:78-114), the two
// entry points py/main.py binds with ctypes (py/main.py:96-109) and calls (:112-126, :151-155): same names, same
// by-value / pointer conventions, so that the reference's unchanged Python front end runs on this library.
//
// The episode loop, the eleven decision networks and the advantage-actor-critic trainer of this framework live in
// Python (fastace_b200/legacy.py: torch for the nets, the CUDA env for the steps).  These C entry points therefore
// call fastace_b200.legacy through the CPython C API of the process they are loaded into — py/main.py's own
// interpreter (ctypes released the GIL for the call; it is re-acquired here) — or, for a caller that is not Python,
// through an interpreter they start themselves.  The C API is resolved at run time (dlsym), so the library has no
// link-time dependency on libpython and loads into any process.
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "../../include/fastace_b200.h"

namespace {

struct PyApi {
    int (*IsInitialized)();
    void (*InitializeEx)(int);
    int (*GILEnsure)();
    void (*GILRelease)(int);
    void* (*ImportModule)(const char*);
    void* (*GetAttrString)(void*, const char*);
    void* (*CallFunction)(void*, const char*, ...);
    void (*DecRef)(void*);
    void (*ErrPrint)();
    int (*RunSimpleString)(const char*);
    bool ok;
};

template <typename T>
bool resolve(void* handle, const char* name, T& fn) {
    fn = reinterpret_cast<T>(dlsym(handle, name));
    return fn != nullptr;
}

PyApi load_python() {
    PyApi a = {};
    void* h = RTLD_DEFAULT;
    if (!dlsym(RTLD_DEFAULT, "Py_IsInitialized")) {
        // not inside a Python process: bring the interpreter's library in (FASTACE_PYTHON_LIB overrides the guesses)
        const char* env = getenv("FASTACE_PYTHON_LIB");
        const char* guesses[] = {env, "libpython3.12.so.1.0", "libpython3.12.so", "libpython3.so", nullptr};
        h = nullptr;
        for (int i = 0; i < 5 && !h; i++) if (guesses[i]) h = dlopen(guesses[i], RTLD_NOW | RTLD_GLOBAL);
        if (!h) { fprintf(stderr, "fastace_b200: no Python C API in this process and libpython could not be loaded (set FASTACE_PYTHON_LIB)\n"); return a; }
    }
    a.ok = resolve(h, "Py_IsInitialized", a.IsInitialized) && resolve(h, "Py_InitializeEx", a.InitializeEx) &&
           resolve(h, "PyGILState_Ensure", a.GILEnsure) && resolve(h, "PyGILState_Release", a.GILRelease) &&
           resolve(h, "PyImport_ImportModule", a.ImportModule) && resolve(h, "PyObject_GetAttrString", a.GetAttrString) &&
           resolve(h, "PyObject_CallFunction", a.CallFunction) && resolve(h, "Py_DecRef", a.DecRef) &&
           resolve(h, "PyErr_Print", a.ErrPrint) && resolve(h, "PyRun_SimpleString", a.RunSimpleString);
    if (!a.ok) fprintf(stderr, "fastace_b200: incomplete Python C API\n");
    return a;
}

// directory that holds the fastace_b200 package = two levels above this shared object
std::string package_parent() {
    Dl_info info;
    if (!dladdr(reinterpret_cast<void*>(&package_parent), &info) || !info.dli_fname) return ".";
    std::string p(info.dli_fname);
    for (int up = 0; up < 2; up++) {
        const size_t k = p.find_last_of('/');
        if (k == std::string::npos) return ".";
        p.erase(k);
    }
    return p.empty() ? "/" : p;
}

// calls fastace_b200.legacy.<fn>(*addresses, flag, value) with the GIL held; returns false on a Python error
bool call_legacy(const char* fn, const void* a0, const void* a1, const void* a2, int flag, double value) {
    static PyApi api = load_python();
    if (!api.ok) return false;
    const bool own_interpreter = !api.IsInitialized();
    if (own_interpreter) api.InitializeEx(0);
    const int gil = api.GILEnsure();
    bool ok = false;
    {
        const std::string boot = "import sys\np = r'''" + package_parent() + "'''\nif p not in sys.path: sys.path.insert(0, p)\n";
        api.RunSimpleString(boot.c_str());
        void* mod = api.ImportModule("fastace_b200.legacy");
        void* f = mod ? api.GetAttrString(mod, fn) : nullptr;
        void* r = f ? api.CallFunction(f, "(KKKid)", (unsigned long long)(uintptr_t)a0, (unsigned long long)(uintptr_t)a1,
                                       (unsigned long long)(uintptr_t)a2, flag, value)
                    : nullptr;
        ok = r != nullptr;
        if (!ok) api.ErrPrint();
        if (r) api.DecRef(r);
        if (f) api.DecRef(f);
        if (mod) api.DecRef(mod);
    }
    api.GILRelease(gil);
    return ok;
}

}  // namespace

extern "C" {

// src/pybindings.cpp:78-89: both structs BY VALUE (SysV: in memory on the caller's stack), nothing returned
void run(fastace_custom_scenario_params_t scenarioParams, fastace_training_params_t trainingParams) {
    call_legacy("_c_run", &scenarioParams, &trainingParams, nullptr, 0, 0.0);
}

// src/pybindings.cpp:92-114: `output` receives trainingParams->numEpisodes losses (caller-owned), the learning rates
// the schedulers end on are written back into *trainingParams.  A failure leaves NaN losses behind.
void train(double* output, const fastace_custom_scenario_params_t* scenarioParams, fastace_training_params_t* trainingParams,
           bool fromPretrained, double perturbationSize) {
    if (!output || !scenarioParams || !trainingParams) return;
    if (!call_legacy("_c_train", output, scenarioParams, trainingParams, fromPretrained ? 1 : 0, perturbationSize))
        for (uint32_t i = 0; i < trainingParams->numEpisodes; i++) output[i] = 0.0 / 0.0;
}

}  // extern "C"
