// Internal interface of hostpack.cpp (host-only helpers of the upload / download pipeline in capi.cu).
#pragma once

typedef void (*vlgp_host_task)(int index, void *arg);

// Runs fn(0..n-1, arg), each index once, on the persistent host pool and the calling thread; returns when all are done.
void vlgp_host_parallel(int n, vlgp_host_task fn, void *arg);

template <class F>
static inline void vlgp_host_parallel_for(int n, F &f) {
    vlgp_host_parallel(n, [](int t, void *a) { (*static_cast<F *>(a))(t); }, &f);
}
