// Does cuTensorMapEncodeTiled accept a zero stride (a "broadcast" dimension)?  Used to decide whether nearest-neighbour
// upsampling can be expressed as TMA addressing of a GEMM operand.   nvcc -o build/tmap_probe tests/probes/tmap_stride0_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
int main() {
    cudaFree(0);
    void* p; cudaMalloc(&p, 1 << 24);
    for (unsigned long long s1 : {0ull, 16ull, 128ull}) {
        CUtensorMap m;
        cuuint64_t dims[4] = {64, 2, 64, 1024}, strides[3] = {s1, 128, 128 * 64};
        cuuint32_t box[4] = {64, 2, 64, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("stride[1] = %llu bytes -> CUresult %d\n", s1, (int)r);
    }
    return 0;
}
