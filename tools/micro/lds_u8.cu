// Micro-benchmark: shared-memory wavefronts of byte loads under different lane address patterns.
// nvcc -arch=sm_100a -o lds_u8 lds_u8.cu ; ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum ./lds_u8
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lds8(unsigned a) { unsigned v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds16(unsigned a) { unsigned v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds32(unsigned a) { unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
template <int MODE>
__global__ void k(unsigned* out, int stride_x100, int rowsplit, int pitch)
{
    __shared__ __align__(128) unsigned char buf[32768];
    for (int i = threadIdx.x; i < 32768; i += blockDim.x) buf[i] = (unsigned char)i;
    __syncthreads();
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    const int lane = threadIdx.x & 31;
    unsigned acc = 0;
    for (int it = 0; it < 64; ++it) {
        int x = (lane * stride_x100) / 100 + it;
        int row = (rowsplit > 0 && lane >= rowsplit) ? 1 : 0;
        unsigned a = base + (unsigned)(row * pitch + x);
        if (MODE == 0) acc += lds8(a);
        if (MODE == 1) acc += lds16(a & ~1u);
        if (MODE == 2) acc += lds32(a & ~3u);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    unsigned* out; cudaMalloc(&out, 1 << 20);
    // name, mode, stride (x100 bytes per lane), rowsplit, pitch
    k<0><<<148, 128>>>(out, 100, 0, 288);   // u8, consecutive bytes
    k<0><<<148, 128>>>(out, 400, 0, 288);   // u8, one word per lane
    k<0><<<148, 128>>>(out, 370, 0, 288);   // u8, stride 3.7 (pair mapping at the C3 centre)
    k<0><<<148, 128>>>(out, 370, 16, 288);  // u8, stride 3.7, lanes 16.. on the next row, pitch 288
    k<0><<<148, 128>>>(out, 370, 16, 256);  // same, pitch 256
    k<0><<<148, 128>>>(out, 740, 0, 288);   // u8, stride 7.4 (4 consecutive columns per lane)
    k<1><<<148, 128>>>(out, 370, 0, 288);   // u16
    k<2><<<148, 128>>>(out, 370, 0, 288);   // u32
    k<2><<<148, 128>>>(out, 400, 0, 288);   // u32 one word per lane
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
