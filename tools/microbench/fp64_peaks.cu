// Microbenchmark: FP64 pipes on B200 (sm_100a). Measures DMMA.8x8x4, DFMA and exp() rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peaks tools/microbench/fp64_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

template<int NACC>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters){
  double a=threadIdx.x*1e-3, b=threadIdx.x*2e-3;
  double c[NACC][2];
  #pragma unroll
  for(int j=0;j<NACC;j++){c[j][0]=0;c[j][1]=0;}
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]),"+d"(c[j][1]) : "d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=c[j][0]+c[j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b){
  double c[NACC];
  #pragma unroll
  for(int j=0;j<NACC;j++) c[j]=threadIdx.x+j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++) c[j]=fma(c[j],a,b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// DMMA and DFMA interleaved in one instruction stream: if the two share the FP64 datapath the time is the SUM of
// the separate times, if they are separate pipes it is the MAX (decides whether CUDA-core FMAs are "free" next to DMMA).
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b){
  double ma=threadIdx.x*1e-3, mb=threadIdx.x*2e-3;
  double c[8][2], f[16];
  #pragma unroll
  for(int j=0;j<8;j++){c[j][0]=0;c[j][1]=0;}
  #pragma unroll
  for(int j=0;j<16;j++) f[j]=threadIdx.x+j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<8;j++){
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]),"+d"(c[j][1]) : "d"(ma),"d"(mb));
      f[2*j]=fma(f[2*j],a,b); f[2*j+1]=fma(f[2*j+1],a,b);
    }
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<8;j++) s+=c[j][0]+c[j][1];
  #pragma unroll
  for(int j=0;j<16;j++) s+=f[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) exp_kernel(double* out, int iters, double a){
  double x0=-(threadIdx.x%97)*0.01, s=0;
  for(int i=0;i<iters;i++){ s+=exp(x0); x0-=a; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  double* out; CK(cudaMalloc(&out, sizeof(double)*148*8*256*4));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for(int ctas_per_sm=1; ctas_per_sm<=4; ctas_per_sm*=2){
    int grid=p.multiProcessorCount*ctas_per_sm; int iters=20000;
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); dmma_kernel<8><<<grid,256>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms,e0,e1);
    }
    double fl=(double)grid*8/*warps*/*iters*8/*acc*/*512.0;
    printf("DMMA.8x8x4 NACC=8 ctas/sm=%d: %.3f ms  %.2f TFLOP/s\n",ctas_per_sm,ms,fl/ms*1e-9);
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); dmma_kernel<24><<<grid,256>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms,e0,e1);
    }
    fl=(double)grid*8*iters*24*512.0;
    printf("DMMA.8x8x4 NACC=24 ctas/sm=%d: %.3f ms  %.2f TFLOP/s\n",ctas_per_sm,ms,fl/ms*1e-9);
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); dfma_kernel<16><<<grid,256>>>(out,iters,1.0000001,1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms,e0,e1);
    }
    fl=(double)grid*256*iters*16*2.0;
    printf("DFMA NACC=16 ctas/sm=%d: %.3f ms  %.2f TFLOP/s\n",ctas_per_sm,ms,fl/ms*1e-9);
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); mixed_kernel<<<grid,256>>>(out,iters,1.0000001,1e-9); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms,e0,e1);
    }
    printf("MIXED 8 DMMA + 16 DFMA per iter ctas/sm=%d: %.3f ms  DMMA part %.2f TFLOP/s + DFMA part %.2f TFLOP/s\n",ctas_per_sm,ms,
           (double)grid*8*iters*8*512.0/ms*1e-9,(double)grid*256*iters*16*2.0/ms*1e-9);
    for(int rep=0;rep<3;rep++){
      cudaEventRecord(e0); exp_kernel<<<grid,256>>>(out,2000,1e-4); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms,e0,e1);
    }
    printf("exp(double) ctas/sm=%d: %.3f ms  %.2f Gexp/s\n",ctas_per_sm,ms,(double)grid*256*2000/ms*1e-6);
  }
  // pinned copy bandwidth
  size_t nb=1ull<<30; void *h,*d; CK(cudaMallocHost(&h,nb)); CK(cudaMalloc(&d,nb));
  for(int rep=0;rep<2;rep++){ cudaEventRecord(e0); cudaMemcpyAsync(d,h,nb,cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
  printf("H2D pinned 1GiB: %.2f GB/s\n",nb/ms*1e-6);
  for(int rep=0;rep<2;rep++){ cudaEventRecord(e0); cudaMemcpyAsync(h,d,nb,cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);}
  printf("D2H pinned 1GiB: %.2f GB/s\n",nb/ms*1e-6);
  return 0;
}
