// Microbenchmark of the DMMA GEMM phase of solve_dmma_kernel<6> (NP=48, LD=52, 12 warps, 1x3 strips):
// variants of the complex product, fragment prefetching and barrier placement.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/gemm_phase_bench tools/gemm_phase_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
constexpr int NP=48, LD=52, PLANE=NP*LD, MAT=2*PLANE, TW=3, NTH=384;

__device__ __forceinline__ void dmma(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
// V=0: 4 real products; V=1: Karatsuba with DADD operand sums; V=2: Karatsuba, sums read from a third smem plane;
// V=3: Karatsuba + explicit double-buffered fragments; V=4: 4 products + explicit double buffering
template<int V>
__device__ __forceinline__ void gemm(double (&acc)[TW][2][2], const double* __restrict__ A, const double* __restrict__ B,
                                     const double* __restrict__ As, const double* __restrict__ Bs, int tr,int tc0,int g,int tg){
  const double* Are=A+(8*tr+g)*LD+tg; const double* Aim=Are+PLANE; const double* Asm=As+(8*tr+g)*LD+tg;
  if (V==0){
    #pragma unroll
    for(int k0=0;k0<NP;k0+=4){ double ar=Are[k0],ai=Aim[k0];
      #pragma unroll
      for(int t=0;t<TW;t++){ const double* bp=B+(k0+tg)*LD+8*(tc0+t)+g; double br=bp[0],bi=bp[PLANE];
        dmma(acc[t][0][0],acc[t][0][1],ar,br); dmma(acc[t][0][0],acc[t][0][1],-ai,bi);
        dmma(acc[t][1][0],acc[t][1][1],ar,bi); dmma(acc[t][1][0],acc[t][1][1],ai,br);} }
  } else if (V==4){
    double ar=Are[0],ai=Aim[0],br[TW],bi[TW];
    #pragma unroll
    for(int t=0;t<TW;t++){ const double* bp=B+tg*LD+8*(tc0+t)+g; br[t]=bp[0]; bi[t]=bp[PLANE]; }
    #pragma unroll
    for(int k0=0;k0<NP;k0+=4){ double nar=0,nai=0,nbr[TW],nbi[TW];
      if(k0+4<NP){ nar=Are[k0+4]; nai=Aim[k0+4];
        #pragma unroll
        for(int t=0;t<TW;t++){ const double* bp=B+(k0+4+tg)*LD+8*(tc0+t)+g; nbr[t]=bp[0]; nbi[t]=bp[PLANE]; } }
      #pragma unroll
      for(int t=0;t<TW;t++){
        dmma(acc[t][0][0],acc[t][0][1],ar,br[t]); dmma(acc[t][1][0],acc[t][1][1],ar,bi[t]); }
      #pragma unroll
      for(int t=0;t<TW;t++){
        dmma(acc[t][0][0],acc[t][0][1],-ai,bi[t]); dmma(acc[t][1][0],acc[t][1][1],ai,br[t]); }
      ar=nar; ai=nai;
      #pragma unroll
      for(int t=0;t<TW;t++){ br[t]=nbr[t]; bi[t]=nbi[t]; } }
  } else {
    double P[TW][3][2];
    #pragma unroll
    for(int t=0;t<TW;t++) for(int i=0;i<3;i++){P[t][i][0]=0;P[t][i][1]=0;}
    if (V==1){
      #pragma unroll
      for(int k0=0;k0<NP;k0+=4){ double ar=Are[k0],ai=Aim[k0],as=ar+ai;
        #pragma unroll
        for(int t=0;t<TW;t++){ const double* bp=B+(k0+tg)*LD+8*(tc0+t)+g; double br=bp[0],bi=bp[PLANE],bs=br+bi;
          dmma(P[t][0][0],P[t][0][1],ar,br); dmma(P[t][1][0],P[t][1][1],ai,bi); dmma(P[t][2][0],P[t][2][1],as,bs);} }
    } else if (V==2){
      #pragma unroll
      for(int k0=0;k0<NP;k0+=4){ double ar=Are[k0],ai=Aim[k0],as=Asm[k0];
        #pragma unroll
        for(int t=0;t<TW;t++){ const double* bp=B+(k0+tg)*LD+8*(tc0+t)+g; double br=bp[0],bi=bp[PLANE],bs=Bs[(k0+tg)*LD+8*(tc0+t)+g];
          dmma(P[t][0][0],P[t][0][1],ar,br); dmma(P[t][1][0],P[t][1][1],ai,bi); dmma(P[t][2][0],P[t][2][1],as,bs);} }
    } else { // V==3
      double ar=Are[0],ai=Aim[0],br[TW],bi[TW];
      #pragma unroll
      for(int t=0;t<TW;t++){ const double* bp=B+tg*LD+8*(tc0+t)+g; br[t]=bp[0]; bi[t]=bp[PLANE]; }
      #pragma unroll
      for(int k0=0;k0<NP;k0+=4){ double nar=0,nai=0,nbr[TW],nbi[TW];
        if(k0+4<NP){ nar=Are[k0+4]; nai=Aim[k0+4];
          #pragma unroll
          for(int t=0;t<TW;t++){ const double* bp=B+(k0+4+tg)*LD+8*(tc0+t)+g; nbr[t]=bp[0]; nbi[t]=bp[PLANE]; } }
        double as=ar+ai;
        #pragma unroll
        for(int t=0;t<TW;t++){ double bs=br[t]+bi[t];
          dmma(P[t][0][0],P[t][0][1],ar,br[t]); dmma(P[t][1][0],P[t][1][1],ai,bi[t]); dmma(P[t][2][0],P[t][2][1],as,bs); }
        ar=nar; ai=nai;
        #pragma unroll
        for(int t=0;t<TW;t++){ br[t]=nbr[t]; bi[t]=nbi[t]; } }
    }
    #pragma unroll
    for(int t=0;t<TW;t++) for(int c=0;c<2;c++){ acc[t][0][c]+=P[t][0][c]-P[t][1][c]; acc[t][1][c]+=P[t][2][c]-P[t][0][c]-P[t][1][c]; }
  }
}

// SYNC: 0 = no barriers between phases, 1 = one __syncthreads per phase, 2 = barrier + write result to T (like the real kernel)
template<int V,int SYNC>
__global__ void __launch_bounds__(NTH,1) phase_kernel(double* out,int reps,long long* clk){
  extern __shared__ __align__(16) double sm[];
  double* Y=sm; double* X=sm+MAT; double* T=sm+2*MAT; double* S1=sm+3*MAT; double* S2=S1+PLANE;
  for(int i=threadIdx.x;i<3*MAT+2*PLANE;i+=NTH) sm[i]=1e-3*((i*37)%101-50);
  __syncthreads();
  int tid=threadIdx.x,w=tid>>5,lane=tid&31,g=lane>>2,tg=lane&3,tr=w/2,tc0=(w%2)*3;
  double K[TW][2][2];
  #pragma unroll
  for(int t=0;t<TW;t++){K[t][0][0]=K[t][0][1]=K[t][1][0]=K[t][1][1]=0;}
  long long t0=clock64();
  for(int r=0;r<reps;r++){
    #pragma unroll 1
    for(int ph=0;ph<9;ph++){
      if(SYNC>=1) __syncthreads();
      if(SYNC==2 && (ph&1)){
        #pragma unroll
        for(int t=0;t<TW;t++){ int rr=8*tr+g,c0=8*(tc0+t)+2*tg;
          *reinterpret_cast<double2*>(T+rr*LD+c0)=make_double2(K[t][0][0]*1e-9,K[t][0][1]*1e-9);
          *reinterpret_cast<double2*>(T+PLANE+rr*LD+c0)=make_double2(K[t][1][0]*1e-9,K[t][1][1]*1e-9);}
        __syncthreads();
      }
      gemm<V>(K,(ph&1)?T:X,Y,S1,S2,tr,tc0,g,tg);
    }
  }
  long long t1=clock64();
  double s=0;
  #pragma unroll
  for(int t=0;t<TW;t++) s+=K[t][0][0]+K[t][0][1]+K[t][1][0]+K[t][1][1];
  out[blockIdx.x*NTH+tid]=s;
  if(tid==0) clk[blockIdx.x]=t1-t0;
}

template<int V,int SYNC> void run(const char* name,double* out,long long* clk,int sms){
  size_t smem=(3*MAT+2*PLANE+8)*sizeof(double);
  CK(cudaFuncSetAttribute(phase_kernel<V,SYNC>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  int reps=200;
  phase_kernel<V,SYNC><<<sms,NTH,smem>>>(out,reps,clk); CK(cudaDeviceSynchronize());
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0)); phase_kernel<V,SYNC><<<sms,NTH,smem>>>(out,reps,clk); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1));
  long long c; CK(cudaMemcpy(&c,clk,8,cudaMemcpyDeviceToHost));
  double per_gemm=(double)c/(reps*9.0);
  // complex 48^3 GEMM = 8*48^3 algorithmic flops
  double tf=8.0*48*48*48*9.0*reps*sms/(ms*1e-3)/1e12;
  printf("%-34s clk/gemm=%8.0f  (ideal 4-mult 6912, 3-mult 5184)  algorithmic TFLOP/s=%6.2f\n",name,per_gemm,tf);
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double* out; long long* clk; CK(cudaMalloc(&out,sizeof(double)*sms*NTH)); CK(cudaMalloc(&clk,8*sms));
  run<0,0>("4mult nosync",out,clk,sms); run<0,1>("4mult sync",out,clk,sms); run<0,2>("4mult sync+Twrite",out,clk,sms);
  run<4,0>("4mult dbuf nosync",out,clk,sms); run<4,2>("4mult dbuf sync+Twrite",out,clk,sms);
  run<1,0>("3mult(DADD) nosync",out,clk,sms); run<1,1>("3mult(DADD) sync",out,clk,sms); run<1,2>("3mult(DADD) sync+Twrite",out,clk,sms);
  run<2,0>("3mult(smem sums) nosync",out,clk,sms); run<2,2>("3mult(smem sums) sync+Twrite",out,clk,sms);
  run<3,0>("3mult dbuf nosync",out,clk,sms); run<3,2>("3mult dbuf sync+Twrite",out,clk,sms);
  return 0;
}
