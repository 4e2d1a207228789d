// Micro-benchmark 2: byte loads whose lanes walk along x (stride s) and along y (row every `rl` lanes),
// for several row pitches.  ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lds8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__global__ void k(unsigned* out, int stride_x100, int lanes_per_row, int pitch, int x0)
{
    __shared__ __align__(128) unsigned char buf[40960];
    for (int i = threadIdx.x; i < 40960; i += blockDim.x) buf[i] = (unsigned char)i;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    const int lane = threadIdx.x & 31;
    unsigned acc = 0;
    for (int it = 0; it < 64; ++it) {
        int x = x0 + (lane * stride_x100) / 100 + (it & 7);
        int row = lane / lanes_per_row + (it >> 3);
        acc += lds8(base + (unsigned)(row * pitch + x));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    unsigned* out; cudaMalloc(&out, 1 << 20);
    const int pitches[4] = {256, 288, 128, 384};
    for (int p = 0; p < 4; ++p)
        for (int lpr = 32; lpr >= 4; lpr /= 2)          // 1, 2, 4, 8 rows across the warp
            k<<<148, 128>>>(out, 369, lpr, pitches[p], 13);
    k<<<148, 128>>>(out, 200, 8, 128, 5);   // narrow piece (mag 1.0), pitch 128
    k<<<148, 128>>>(out, 120, 8, 128, 5);   // edge piece (mag 0.6)
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
