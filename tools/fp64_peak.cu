// FP64 peak microbenchmark for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate, and a
// smem-fed DMMA loop.  Produces the "tensor" roofline denominator that MEASURED_PEAKS.json lacks.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(1024) dmma_chain(double* out, int iters){
  double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
  double c[NACC][2];
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=0;c[i][1]=0;}
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   :"+d"(c[i][0]),"+d"(c[i][1]):"d"(a),"d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void __launch_bounds__(1024) dfma_chain(double* out, int iters){
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-7;
  double c[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) c[i]=i;
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int i=0;i<NACC;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// smem-fed: each warp computes a (8*MT)x(8*NT) real tile over K=48 from smem, repeated.
template<int MT,int NT>
__global__ void __launch_bounds__(512) dmma_smem(double* out, int iters){
  __shared__ double As[48*52];
  __shared__ double Bs[48*52];
  for(int i=threadIdx.x;i<48*52;i+=blockDim.x){As[i]=i*1e-4;Bs[i]=1.0-i*1e-5;}
  __syncthreads();
  int lane=threadIdx.x&31, g=lane>>2, tg=lane&3;
  double c[MT][NT][2];
  #pragma unroll
  for(int i=0;i<MT;i++)
  #pragma unroll
  for(int j=0;j<NT;j++){c[i][j][0]=0;c[i][j][1]=0;}
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int k=0;k<48;k+=4){
      double a[MT],b[NT];
      #pragma unroll
      for(int i=0;i<MT;i++) a[i]=As[(i*8+g)*52+k+tg];
      #pragma unroll
      for(int j=0;j<NT;j++) b[j]=Bs[(j*8+g)*52+k+tg];
      #pragma unroll
      for(int i=0;i<MT;i++)
      #pragma unroll
      for(int j=0;j<NT;j++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   :"+d"(c[i][j][0]),"+d"(c[i][j][1]):"d"(a[i]),"d"(b[j]));
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<MT;i++)
  #pragma unroll
  for(int j=0;j<NT;j++) s+=c[i][j][0]+c[i][j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f,int reps){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
  return best;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*1024));
  printf("{\"gpu\":\"%s\",\"sms\":%d", p.name, sms);
  const int iters=20000;
  // DMMA register-only chains, varying warps/SM and accumulators
  {
    int wps[]={4,8,16,32};
    for(int wi=0;wi<4;wi++){
      int threads=wps[wi]*32;
      float ms=timeit([&]{dmma_chain<8><<<sms,threads>>>(out,iters);},5);
      double flops=2.0*256*8*(double)iters*wps[wi]*sms;
      printf(",\"dmma_reg_w%d_tflops\":%.2f", wps[wi], flops/ms*1e-9);
    }
    float ms=timeit([&]{dmma_chain<2><<<sms,512>>>(out,iters);},5);
    printf(",\"dmma_reg_w16_acc2_tflops\":%.2f", 2.0*256*2*(double)iters*16*sms/ms*1e-9);
    ms=timeit([&]{dmma_chain<1><<<sms,128>>>(out,iters);},5);
    printf(",\"dmma_latency_clk_est_ns\":%.2f", ms*1e6/iters);
  }
  {
    int wps[]={8,16,32};
    for(int wi=0;wi<3;wi++){
      int threads=wps[wi]*32;
      float ms=timeit([&]{dfma_chain<16><<<sms*2,threads>>>(out,iters);},5);
      double flops=2.0*16*(double)iters*threads*sms*2;
      printf(",\"dfma_w%d_tflops\":%.2f", wps[wi]*2, flops/ms*1e-9);
    }
  }
  {
    float ms=timeit([&]{dmma_smem<1,3><<<sms,384>>>(out,2000);},5);
    printf(",\"dmma_smem_1x3_w12_tflops\":%.2f", 2.0*256*3*12*2000.0*12*sms/ms*1e-9);
    ms=timeit([&]{dmma_smem<2,2><<<sms,512>>>(out,2000);},5);
    printf(",\"dmma_smem_2x2_w16_tflops\":%.2f", 2.0*256*4*12*2000.0*16*sms/ms*1e-9);
    ms=timeit([&]{dmma_smem<3,3><<<sms,128>>>(out,2000);},5);
    printf(",\"dmma_smem_3x3_w4_tflops\":%.2f", 2.0*256*9*12*2000.0*4*sms/ms*1e-9);
    ms=timeit([&]{dmma_smem<3,3><<<sms,256>>>(out,2000);},5);
    printf(",\"dmma_smem_3x3_w8_tflops\":%.2f", 2.0*256*9*12*2000.0*8*sms/ms*1e-9);
    ms=timeit([&]{dmma_smem<2,3><<<sms,384>>>(out,2000);},5);
    printf(",\"dmma_smem_2x3_w12_tflops\":%.2f", 2.0*256*6*12*2000.0*12*sms/ms*1e-9);
  }
  // sustained: run dmma for ~3 s and report
  {
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n=0; float total=0;
    while(total<3000.f){ for(int i=0;i<10;i++) dmma_chain<8><<<sms,512>>>(out,iters); n+=10; CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&total,e0,e1)); }
    printf(",\"dmma_sustained_tflops\":%.2f", 2.0*256*8*(double)iters*16*sms*n/total*1e-9);
  }
  printf("}\n");
  return 0;
}
