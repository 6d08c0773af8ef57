// FP64 pipe microbenchmark for B200: DMMA m8n8k4 / m16n8k16 vs DFMA, register-resident.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ void dmma884(double &c0,double &c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4],const double (&a)[8],const double (&b)[4]){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3])
   :"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
template<int NACC>
__global__ void k_dmma884(double* out,int iters){
  double c0[NACC],c1[NACC]; double a=threadIdx.x*1e-3, b=1.0+threadIdx.x*1e-6;
  #pragma unroll
  for(int i=0;i<NACC;i++){c0[i]=i;c1[i]=-i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c0[i],c1[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c0[i]+c1[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void k_dmma16816(double* out,int iters){
  double c[NACC][4]; double a[8],b[4];
  #pragma unroll
  for(int i=0;i<8;i++) a[i]=threadIdx.x*1e-3+i;
  #pragma unroll
  for(int i=0;i<4;i++) b[i]=1.0+threadIdx.x*1e-6*i;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma16816(c[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void k_dfma(double* out,int iters){
  double c[NACC]; double a=1.0+threadIdx.x*1e-9, b=threadIdx.x*1e-6;
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best=1e30f;
  for(int r=0;r<5;r++){ cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n",p.name,p.multiProcessorCount,p.clockRate);
  double* out; CK(cudaMalloc(&out,sizeof(double)*148*8*1024));
  int nsm=p.multiProcessorCount;
  const int iters=20000;
  for(int warps: {4,8,16,32}){
    for(int bps: {1,2}){
      if(warps*bps>64) continue;
      int thr=warps*32; int grid=nsm*bps;
      float ms=timeit([&]{k_dmma884<16><<<grid,thr>>>(out,iters);});
      double fl=(double)grid*warps*iters*16*512.0;
      printf("DMMA884   acc16 warps/cta %2d cta/sm %d : %.3f ms  %.2f TFLOP/s\n",warps,bps,ms,fl/ms/1e9);
      ms=timeit([&]{k_dmma884<32><<<grid,thr>>>(out,iters/2);});
      fl=(double)grid*warps*(iters/2)*32*512.0;
      printf("DMMA884   acc32 warps/cta %2d cta/sm %d : %.3f ms  %.2f TFLOP/s\n",warps,bps,ms,fl/ms/1e9);
      ms=timeit([&]{k_dmma16816<8><<<grid,thr>>>(out,iters/8);});
      fl=(double)grid*warps*(iters/8)*8*(2.0*16*8*16);
      printf("DMMA16816 acc8  warps/cta %2d cta/sm %d : %.3f ms  %.2f TFLOP/s\n",warps,bps,ms,fl/ms/1e9);
      ms=timeit([&]{k_dfma<16><<<grid,thr>>>(out,iters*4);});
      fl=(double)grid*thr*(iters*4.0)*16*2.0;
      printf("DFMA      acc16 warps/cta %2d cta/sm %d : %.3f ms  %.2f TFLOP/s\n",warps,bps,ms,fl/ms/1e9);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
