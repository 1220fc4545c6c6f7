#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned a, unsigned b, int* out){ bool ph, pl; unsigned r = __vibmax_s16x2(a, b, &ph, &pl); out[0]=r; out[1]=ph; out[2]=pl; }
int main(){ int* d; cudaMalloc(&d, 12); int h[3];
  unsigned tests[][2] = {{0x00050003u,0x00050004u},{0x00060003u,0x00050003u},{0x00040003u,0x00050002u}};
  for (auto& t: tests){ k<<<1,1>>>(t[0],t[1],d); cudaMemcpy(h,d,12,cudaMemcpyDeviceToHost); printf("a=%08x b=%08x -> r=%08x pred_hi=%d pred_lo=%d\n", t[0],t[1],h[0],h[1],h[2]); }
  return 0; }
