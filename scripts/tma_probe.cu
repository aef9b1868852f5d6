// Standalone probe: which cuTensorMapEncodeTiled configurations load correctly on this GPU?
// usage: tma_probe <elem 1|4> <boxw> <boxh> <l2promo 0..3> <w> <h> <pitch_elems> <x> <y>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "../miniengineao_b200/csrc/common.cuh"
using namespace meao;
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int bytes, unsigned char *out)
{
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, bytes); tma_load_2d(buf, &map, x, y, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char **argv)
{
    int elem = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), l2 = atoi(argv[4]);
    int w = atoi(argv[5]), h = atoi(argv[6]), pitch = atoi(argv[7]), x = atoi(argv[8]), y = atoi(argv[9]);
    void *fn; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) { printf("no entry point\n"); return 2; }
    size_t bytes = (size_t)pitch * h * elem;
    unsigned char *g; cudaMalloc(&g, bytes);
    std::vector<unsigned char> hbuf(bytes);
    for (size_t i = 0; i < bytes; i++) hbuf[i] = (unsigned char)(i * 131 + 7);
    cudaMemcpy(g, hbuf.data(), bytes, cudaMemcpyHostToDevice);
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h}; cuuint64_t strides[1] = {(cuuint64_t)pitch * elem};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((PFN_encodeTiled)fn)(&m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d ", (int)r);
    if (r) { printf("\n"); return 1; }
    int tb = bw * bh * elem;
    unsigned char *out; cudaMalloc(&out, tb);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
    probe<<<1, 128, tb + 128>>>(m, x, y, tb, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run: %s ", cudaGetErrorString(e));
    if (e) { printf("\n"); return 1; }
    std::vector<unsigned char> o(tb);
    cudaMemcpy(o.data(), out, tb, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < bh; r2++) for (int c = 0; c < bw * elem; c++) {
        int gx = x * elem + c, gy = y + r2;
        unsigned char exp = (gx >= 0 && gx < w * elem && gy >= 0 && gy < h) ? hbuf[(size_t)gy * pitch * elem + gx] : 0;
        if (o[r2 * bw * elem + c] != exp) bad++;
    }
    printf("bad bytes=%d\n", bad);
    return 0;
}
