// Micro-benchmark: texture-unit fetch rate on a pitch-linear 8-bit 4K image for the access patterns
// the warp would generate.  nvcc -arch=sm_100a -O3 -o tex_rate tex_rate.cu ; ./tex_rate
// Each warp walks a band of the image: lane l fetches at x = x0 + l * stride (+ a per-lane vertical
// slope), 64 fetches per thread per launch round, rows advancing by `dy` per fetch.
#include <cstdio>
#include <cuda_runtime.h>
enum { FILTER = 0, GATHER = 1, POINT = 2 };
template <int MODE, typename T>
__global__ void __launch_bounds__(128) k(cudaTextureObject_t tex, float stride, float slope, float dy, int W, int H,
                                         float* out, int iters)
{
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // every warp gets its own band origin (pseudo-random but L2/DRAM friendly: consecutive warps are neighbours)
    const int bands_x = (int)(W / (32 * stride)) > 0 ? (int)(W / (32 * stride)) : 1;
    const float x0 = (warp % bands_x) * 32 * stride + 0.37f;
    float y = ((warp / bands_x) * 9) % (H - 80) + 0.41f;
    const float x = x0 + lane * stride, ys = lane * slope;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        if (MODE == GATHER) {
            float4 v = tex2Dgather<float4>(tex, x, y + ys, 0);
            acc += v.x + v.y + v.z + v.w;
        } else if (sizeof(T) == 1) {
            acc += tex2D<float>(tex, x, y + ys);
        } else {
            float2 v = tex2D<float2>(tex, x, y + ys);
            acc += v.x + v.y;
        }
        y += dy;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE, typename T>
void run(const char* name, void* img, size_t pitch, int W, int H, float stride, float slope, float dy, float* out)
{
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = img;
    rd.res.pitch2D.pitchInBytes = pitch; rd.res.pitch2D.width = W / sizeof(T); rd.res.pitch2D.height = H;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<T>();
    cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = MODE == FILTER ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { printf("%s: texture failed\n", name); cudaGetLastError(); return; }
    const int blocks = 148 * 16, iters = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE, T><<<blocks, 128>>>(tex, stride, slope, dy, W / (int)sizeof(T), H, out, iters);
    cudaEventRecord(a);
    for (int r = 0; r < 10; ++r) k<MODE, T><<<blocks, 128>>>(tex, stride, slope, dy, W / (int)sizeof(T), H, out, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double fetches = 10.0 * blocks * 128 * iters;
    printf("%-44s stride %.2f slope %.3f dy %.2f : %7.1f Gfetch/s  (%.2f per clk per SM at 1.965 GHz)  %s\n", name, stride, slope, dy,
           fetches / ms * 1e-6, fetches / ms * 1e-6 / 148 / 1.965, cudaGetErrorString(cudaGetLastError()));
    cudaDestroyTextureObject(tex);
}
int main()
{
    const int W = 3840, H = 2160 * 3 / 2 * 16;
    unsigned char* img; cudaMalloc(&img, (size_t)W * H); cudaMemset(img, 77, (size_t)W * H);
    float* out; cudaMalloc(&out, 148 * 16 * 128 * 4);
    run<FILTER, unsigned char>("u8 bilinear, lanes adjacent", img, W, W, H, 1.0f, 0.f, 1.0f, out);
    run<FILTER, unsigned char>("u8 bilinear, lanes 1.84 px apart", img, W, W, H, 1.84f, 0.f, 1.84f, out);
    run<FILTER, unsigned char>("u8 bilinear, lanes 3.7 px apart (pair map)", img, W, W, H, 3.7f, 0.f, 1.84f, out);
    run<FILTER, unsigned char>("u8 bilinear, 3.7 px apart, sloped", img, W, W, H, 3.7f, 0.05f, 1.84f, out);
    run<FILTER, unsigned char>("u8 bilinear, lanes 7.4 px apart", img, W, W, H, 7.4f, 0.f, 1.84f, out);
    run<FILTER, unsigned char>("u8 bilinear, same texel all lanes", img, W, W, H, 0.0f, 0.f, 0.0f, out);
    run<FILTER, unsigned char>("u8 bilinear, adjacent, no row advance", img, W, W, H, 1.0f, 0.f, 0.0f, out);
    run<POINT, unsigned char>("u8 point, lanes 1.84 px apart", img, W, W, H, 1.84f, 0.f, 1.84f, out);
    run<POINT, unsigned char>("u8 point, lanes 3.7 px apart", img, W, W, H, 3.7f, 0.f, 1.84f, out);
    run<GATHER, unsigned char>("u8 gather4, lanes 1.84 px apart", img, W, W, H, 1.84f, 0.f, 1.84f, out);
    run<GATHER, unsigned char>("u8 gather4, lanes 3.7 px apart", img, W, W, H, 3.7f, 0.f, 1.84f, out);
    run<FILTER, uchar2>("u8x2 bilinear, lanes 1.84 texels apart", img, W, W, H, 1.84f, 0.f, 1.84f, out);
    run<FILTER, uchar2>("u8x2 bilinear, lanes 3.7 texels apart", img, W, W, H, 3.7f, 0.f, 1.84f, out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
