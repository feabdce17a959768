// nezha_b200/csrc/nz_cufile.cu — fill from the PV through cuFile / GPUDirect Storage.
//
// BASELINE config #5 ("cold fill from NVMe PV via cuFile/GDS"); replaces the reference's
// "proxy can run on a PV that has all the needed dataset beforehand" (README.md:123) cold
// start.  libcufile is loaded with dlopen so libnezha_gpu.so has no hard dependency on it;
// when the library, the driver or the file system cannot do it the caller falls back to
// pread + pinned bounce buffers (nz_api.cu), and nz_stats says which path carried the bytes.
#include <cuda_runtime.h>
#include <cufile.h>
#include <dlfcn.h>
#include <errno.h>
#include <string.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <mutex>

#include "nz_internal.h"

namespace {
struct Api {
    void* lib = nullptr;
    CUfileError_t (*DriverOpen)() = nullptr;
    CUfileError_t (*HandleRegister)(CUfileHandle_t*, CUfileDescr_t*) = nullptr;
    void (*HandleDeregister)(CUfileHandle_t) = nullptr;
    ssize_t (*Read)(CUfileHandle_t, void*, size_t, off_t, off_t) = nullptr;
    bool ok = false;
};
Api g_api;
std::once_flag g_once;
// why the last attempt (or the driver open) did not carry the bytes: read back through nz_cufile_status
std::mutex g_why_mu;
char g_why[256] = "not tried";
void why(const char* stage, int err, int cu_err) {
    std::lock_guard<std::mutex> g(g_why_mu);
    snprintf(g_why, sizeof g_why, "%s: cufile err %d, cuda err %d, errno %d", stage, err, cu_err, errno);
}

void init_api() {
    if (const char* ev = getenv("NZ_CUFILE"))
        if (atoi(ev) == 0) return;
    const char* names[] = {"libcufile.so.0", "libcufile.so", "/usr/local/cuda/lib64/libcufile.so.0"};
    for (const char* n : names) {
        g_api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) { why("dlopen libcufile", 0, 0); return; }
    g_api.DriverOpen = (CUfileError_t(*)())dlsym(g_api.lib, "cuFileDriverOpen");
    g_api.HandleRegister = (CUfileError_t(*)(CUfileHandle_t*, CUfileDescr_t*))dlsym(g_api.lib, "cuFileHandleRegister");
    g_api.HandleDeregister = (void (*)(CUfileHandle_t))dlsym(g_api.lib, "cuFileHandleDeregister");
    g_api.Read = (ssize_t(*)(CUfileHandle_t, void*, size_t, off_t, off_t))dlsym(g_api.lib, "cuFileRead");
    if (!g_api.DriverOpen || !g_api.HandleRegister || !g_api.HandleDeregister || !g_api.Read) { why("dlsym", 0, 0); return; }
    const CUfileError_t st = g_api.DriverOpen();
    g_api.ok = st.err == CU_FILE_SUCCESS;
    if (!g_api.ok) why("cuFileDriverOpen", (int)st.err, (int)st.cu_err);
    else why("driver open", 0, 0);
}
}  // namespace

// Reads [file_off, file_off+len) of fd into dev_base+dev_off.  Returns bytes read through cuFile
// (== len on success), or -1 when cuFile is unavailable / refused this file (caller falls back).
int64_t nz_cufile_read(int fd, uint64_t file_off, void* dev_base, uint64_t dev_off, uint64_t len) {
    std::call_once(g_once, init_api);
    if (!g_api.ok) return -1;
    // GDS wants O_DIRECT; reopen the same file that way, else use the descriptor as given (compat mode)
    char path[64];
    snprintf(path, sizeof path, "/proc/self/fd/%d", fd);
    int dfd = open(path, O_RDONLY | O_DIRECT);
    const int use = dfd >= 0 ? dfd : fd;
    CUfileDescr_t d;
    memset(&d, 0, sizeof d);
    d.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;
    d.handle.fd = use;
    CUfileHandle_t h;
    CUfileError_t st = g_api.HandleRegister(&h, &d);
    if (st.err != CU_FILE_SUCCESS && dfd >= 0) {          // the O_DIRECT descriptor was refused: the plain one (compat mode)
        why("cuFileHandleRegister(O_DIRECT)", (int)st.err, (int)st.cu_err);
        close(dfd);
        dfd = -1;
        d.handle.fd = fd;
        st = g_api.HandleRegister(&h, &d);
    }
    if (st.err != CU_FILE_SUCCESS) {
        why("cuFileHandleRegister", (int)st.err, (int)st.cu_err);
        if (dfd >= 0) close(dfd);
        return -1;
    }
    uint64_t done = 0;
    int64_t rc = 0;
    while (done < len) {
        const size_t want = (size_t)((len - done) > (64ull << 20) ? (64ull << 20) : (len - done));
        const ssize_t r = g_api.Read(h, dev_base, want, (off_t)(file_off + done), (off_t)(dev_off + done));
        if (r <= 0) { why("cuFileRead", (int)r, 0); rc = -1; break; }
        done += (uint64_t)r;
    }
    g_api.HandleDeregister(h);
    if (dfd >= 0) close(dfd);
    if (rc >= 0) why("ok", 0, 0);
    return rc < 0 ? -1 : (int64_t)done;
}

void nz_cufile_why(char* out, size_t cap) {
    std::lock_guard<std::mutex> g(g_why_mu);
    snprintf(out, cap, "%s", g_why);
}
