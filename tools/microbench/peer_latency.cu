// Peer-memory signalling costs between two B200s over NVLink (one process, two devices, peer access enabled):
//   (1) __threadfence_system() with nothing outstanding, (2) after one peer store, (3) after 64 peer stores,
//   (4) flag ping-pong round trip (st.volatile to the peer, ld.volatile poll locally).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_latency peer_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) { unsigned long long v; asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

__global__ void k_fences(double* peer_buf, double* local_buf, unsigned long long* out) {
    const int reps = 200;
    unsigned long long t0 = gtime();
    for (int i = 0; i < reps; ++i) __threadfence_system();
    unsigned long long t1 = gtime();
    for (int i = 0; i < reps; ++i) { peer_buf[i] = i; __threadfence_system(); }
    unsigned long long t2 = gtime();
    for (int i = 0; i < reps; ++i) { for (int j = 0; j < 64; ++j) peer_buf[1024 + j * 16] = i; __threadfence_system(); }
    unsigned long long t3 = gtime();
    for (int i = 0; i < reps; ++i) { local_buf[i] = i; __threadfence(); }
    unsigned long long t4 = gtime();
    for (int i = 0; i < reps; ++i) { peer_buf[i] = i; __threadfence(); }
    unsigned long long t5 = gtime();
    out[0] = (t1 - t0) / reps; out[1] = (t2 - t1) / reps; out[2] = (t3 - t2) / reps; out[3] = (t4 - t3) / reps; out[4] = (t5 - t4) / reps;
}

// ping-pong: rank 0 sends i, rank 1 answers i
__global__ void k_pingpong(int me, unsigned long long* my_flag, unsigned long long* peer_flag, int reps, unsigned long long* out) {
    unsigned long long t0 = gtime();
    for (int i = 1; i <= reps; ++i) {
        if (me == 0) {
            st_flag(peer_flag, i);
            unsigned spins = 0;
            while (ld_flag(my_flag) < (unsigned long long)i) if (++spins > (1u << 26)) { out[1] = 1; return; }
        } else {
            unsigned spins = 0;
            while (ld_flag(my_flag) < (unsigned long long)i) if (++spins > (1u << 26)) { out[1] = 1; return; }
            st_flag(peer_flag, i);
        }
    }
    out[0] = (gtime() - t0) / reps;
}

int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    double *buf[2]; unsigned long long *flag[2], *out[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&buf[d], 1 << 20)); CK(cudaMalloc(&flag[d], 256)); CK(cudaMalloc(&out[d], 256));
        CK(cudaMemset(flag[d], 0, 256)); CK(cudaMemset(out[d], 0, 256));
    }
    unsigned long long h[8];
    CK(cudaSetDevice(0));
    k_fences<<<1, 1>>>(buf[1], buf[0], out[0]);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out[0], 64, cudaMemcpyDeviceToHost));
    printf("fence.sys idle %llu ns | 1 peer store + fence.sys %llu ns | 64 peer stores + fence.sys %llu ns | local store + fence.gpu %llu ns | peer store + fence.gpu %llu ns\n",
           h[0], h[1], h[2], h[3], h[4]);
    cudaStream_t s[2];
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaStreamCreate(&s[d])); }
    CK(cudaSetDevice(1)); k_pingpong<<<1, 1, 0, s[1]>>>(1, flag[1], flag[0], 1000, out[1]);
    CK(cudaSetDevice(0)); k_pingpong<<<1, 1, 0, s[0]>>>(0, flag[0], flag[1], 1000, out[0]);
    for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, out[0], 64, cudaMemcpyDeviceToHost));
    printf("flag ping-pong round trip %llu ns (timeout flag %llu)\n", h[0], h[1]);
    return 0;
}
