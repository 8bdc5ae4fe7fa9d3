// Scratch timing probe for the SPD-inverse building blocks (not product code).
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <cstdio>
#include <vector>
#include <functional>
#define CK(x) do{auto e=(x); if(e!=0){printf("ERR %d line %d\n",(int)e,__LINE__); return 1;}}while(0)
__global__ void fill_spd(double* A, int n, size_t ld){ size_t i=blockIdx.x*blockDim.x+threadIdx.x; if(i>=(size_t)n*n) return; int r=i%n,c=i/n; A[(size_t)c*ld+r]=(r==c)?(double)n:1.0/(1.0+abs(r-c)); }
int main(int argc,char**argv){
  int n= argc>1?atoi(argv[1]):10000; size_t ld=n;
  cudaStream_t s; cudaStreamCreate(&s);
  cublasHandle_t bh; cublasCreate(&bh); cublasSetStream(bh,s);
  cusolverDnHandle_t sh; cusolverDnCreate(&sh); cusolverDnSetStream(sh,s);
  double *A,*B,*C; cudaMalloc(&A,ld*n*8); cudaMalloc(&B,ld*n*8); cudaMalloc(&C,ld*n*8);
  int *info; cudaMalloc(&info,4);
  int lw1,lw2; CK(cusolverDnDpotrf_bufferSize(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,&lw1)); CK(cusolverDnDpotri_bufferSize(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,&lw2));
  size_t wd=0,wh=0; CK(cusolverDnXtrtri_bufferSize(sh,CUBLAS_FILL_MODE_UPPER,CUBLAS_DIAG_NON_UNIT,n,CUDA_R_64F,A,ld,&wd,&wh));
  int lw=lw1>lw2?lw1:lw2; double* w; cudaMalloc(&w,(size_t)lw*8+wd+1024); std::vector<char> hw(wh+1);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit=[&](const char* name,double flops,std::function<void()> prep,std::function<int()> f){
    float best=1e30; for(int r=0;r<3;r++){ prep(); cudaStreamSynchronize(s); cudaEventRecord(e0,s); int rc=f(); cudaEventRecord(e1,s); cudaStreamSynchronize(s); if(rc){printf("%s rc=%d\n",name,rc);} float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;}
    printf("n=%d %-18s %8.2f ms  %6.2f TFLOP/s\n",n,name,best,flops/best/1e9); fflush(stdout); };
  double n3=(double)n*n*n; const double one=1,zero=0,mone=-1;
  auto fill=[&](){ fill_spd<<<(unsigned)(((size_t)n*n+255)/256),256,0,s>>>(A,n,ld); };
  auto nop=[&](){};
  timeit("potrf",n3/3,fill,[&](){return (int)cusolverDnDpotrf(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,w,lw,info);});
  timeit("potri(after potrf)",2*n3/3,[&](){fill(); cusolverDnDpotrf(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,w,lw,info);},[&](){return (int)cusolverDnDpotri(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,w,lw,info);});
  timeit("Xtrtri",n3/3,[&](){fill(); cusolverDnDpotrf(sh,CUBLAS_FILL_MODE_UPPER,n,A,ld,w,lw,info);},[&](){return (int)cusolverDnXtrtri(sh,CUBLAS_FILL_MODE_UPPER,CUBLAS_DIAG_NON_UNIT,n,CUDA_R_64F,A,ld,w,wd,hw.data(),wh,info);});
  timeit("dgemm n^3",2*n3,nop,[&](){return (int)cublasDgemm(bh,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,ld,B,ld,&zero,C,ld);});
  timeit("dgemm TN",2*n3,nop,[&](){return (int)cublasDgemm(bh,CUBLAS_OP_T,CUBLAS_OP_N,n,n,n,&one,A,ld,B,ld,&zero,C,ld);});
  timeit("dsyrk n^3",n3,nop,[&](){return (int)cublasDsyrk(bh,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_N,n,n,&one,A,ld,&zero,C,ld);});
  timeit("dtrmm (oop)",n3,nop,[&](){return (int)cublasDtrmm(bh,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_N,CUBLAS_DIAG_NON_UNIT,n,n,&one,A,ld,B,ld,C,ld);});
  timeit("dtrsm n rhs",n3,[&](){fill(); cudaMemsetAsync(B,0,ld*n*8,s);},[&](){return (int)cublasDtrsm(bh,CUBLAS_SIDE_LEFT,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_N,CUBLAS_DIAG_NON_UNIT,n,n,&one,A,ld,B,ld);});
  // half-size gemms (the recursive inverse's top level)
  int h=n/2; timeit("dgemm (n/2)^3",2.0*h*h*h,nop,[&](){return (int)cublasDgemm(bh,CUBLAS_OP_N,CUBLAS_OP_N,h,h,h,&one,A,ld,B,ld,&zero,C,ld);});
  timeit("dsyrk 160 (gram)",(double)n*n*160,nop,[&](){return (int)cublasDsyrk(bh,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_T,n,160,&one,A,160,&zero,C,ld);});
  // small potrf+potri for base cases
  for(int b: {128,256,512,1024}){ int l1,l2; cusolverDnDpotrf_bufferSize(sh,CUBLAS_FILL_MODE_UPPER,b,A,ld,&l1); cusolverDnDpotri_bufferSize(sh,CUBLAS_FILL_MODE_UPPER,b,A,ld,&l2); char nm[64]; sprintf(nm,"potrf+potri b=%d",b);
    timeit(nm,(double)b*b*b,fill,[&](){ cusolverDnDpotrf(sh,CUBLAS_FILL_MODE_UPPER,b,A,ld,w,lw,info); return (int)cusolverDnDpotri(sh,CUBLAS_FILL_MODE_UPPER,b,A,ld,w,lw,info);}); }
  // SVD pieces: geqrf on n x 160, gesvdj 160x160, gesvd n x 160
  { int m=n, k=160; double* tau; cudaMalloc(&tau,k*8); int lq; cusolverDnDgeqrf_bufferSize(sh,m,k,A,m,&lq); double* wq; cudaMalloc(&wq,(size_t)lq*8);
    timeit("geqrf n x 160",2.0*m*k*k,fill,[&](){return (int)cusolverDnDgeqrf(sh,m,k,A,m,tau,wq,lq,info);});
    gesvdjInfo_t gi; cusolverDnCreateGesvdjInfo(&gi); double *S,*U,*V; cudaMalloc(&S,k*8); cudaMalloc(&U,k*k*8); cudaMalloc(&V,k*k*8); int lj; cusolverDnDgesvdj_bufferSize(sh,CUSOLVER_EIG_MODE_VECTOR,1,k,k,B,k,S,U,k,V,k,&lj,gi); double* wj; cudaMalloc(&wj,(size_t)lj*8);
    timeit("gesvdj 160x160",1,[&](){fill_spd<<<(k*k+255)/256,256,0,s>>>(B,k,k);},[&](){return (int)cusolverDnDgesvdj(sh,CUSOLVER_EIG_MODE_VECTOR,1,k,k,B,k,S,U,k,V,k,wj,lj,info,gi);});
    int lg; cusolverDnDgesvd_bufferSize(sh,k,k,&lg); double* wg; cudaMalloc(&wg,(size_t)lg*8);
    timeit("gesvd 160x160",1,[&](){fill_spd<<<(k*k+255)/256,256,0,s>>>(B,k,k);},[&](){return (int)cusolverDnDgesvd(sh,'S','S',k,k,B,k,S,U,k,V,k,wg,lg,nullptr,info);});
    int ls; cusolverDnDsyevd_bufferSize(sh,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_UPPER,k,B,k,S,&ls); double* wsy; cudaMalloc(&wsy,(size_t)ls*8);
    timeit("syevd 160x160",1,[&](){fill_spd<<<(k*k+255)/256,256,0,s>>>(B,k,k);},[&](){return (int)cusolverDnDsyevd(sh,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_UPPER,k,B,k,S,wsy,ls,info);});
  }
  if(n<=4096){ int ls; double* S; cudaMalloc(&S,n*8); cusolverDnDsyevd_bufferSize(sh,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_UPPER,n,A,ld,S,&ls); double* wsy; cudaMalloc(&wsy,(size_t)ls*8);
    timeit("syevd n",1,fill,[&](){return (int)cusolverDnDsyevd(sh,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_UPPER,n,A,ld,S,wsy,ls,info);}); }
  int lg; CK(cusolverDnDgetrf_bufferSize(sh,n,n,A,ld,&lg)); double* wl; cudaMalloc(&wl,(size_t)lg*8); int* piv; cudaMalloc(&piv,n*4);
  timeit("getrf",2*n3/3,fill,[&](){return (int)cusolverDnDgetrf(sh,n,n,A,ld,wl,piv,info);});
  timeit("getrs n rhs",2*n3,[&](){fill(); cusolverDnDgetrf(sh,n,n,A,ld,wl,piv,info); cudaMemsetAsync(B,0,ld*n*8,s);},[&](){return (int)cusolverDnDgetrs(sh,CUBLAS_OP_N,n,n,A,ld,piv,B,ld,info);});
  return 0; }
