// ipc_probe.cu — CUDA IPC peer-pull sanity check under the MPI shim (2 ranks, 2 GPUs).
// Variants: data written by a kernel BEFORE the peer opens the mapping, AFTER it (what the SUMMA
// stores do: export at create, fill later), and after with an H2D copy instead of a kernel.
#include <cuda_runtime.h>
#include <mpi.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/time.h>
#include <vector>
#define CK(x) do { cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"[%d] %s failed: %s line %d\n",rank,#x,cudaGetErrorString(e),__LINE__); MPI_Abort(MPI_COMM_WORLD,1);} } while(0)
static double now(){ struct timeval tv; gettimeofday(&tv,0); return tv.tv_sec+tv.tv_usec*1e-6; }
__global__ void fillk(double* p, size_t n, double v){ for(size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x;i<n;i+=(size_t)gridDim.x*blockDim.x) p[i]=v+i; }
int main(int argc,char**argv){
  int rank,size; MPI_Init(&argc,&argv); MPI_Comm_rank(MPI_COMM_WORLD,&rank); MPI_Comm_size(MPI_COMM_WORLD,&size);
  CK(cudaSetDevice(rank)); CK(cudaFree(0));
  cudaStream_t s, s2; CK(cudaStreamCreateWithFlags(&s,cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2,cudaStreamNonBlocking));
  size_t bytes=(size_t)2<<20, n=bytes/8;
  for(int variant=0; variant<4; ++variant){
    double *mine; CK(cudaMalloc(&mine,bytes)); CK(cudaMemset(mine,0,bytes));
    if(variant==0){ fillk<<<64,256,0,s2>>>(mine,n,1000.0*(rank+1)); CK(cudaStreamSynchronize(s2)); }
    cudaIpcMemHandle_t h[2]; CK(cudaIpcGetMemHandle(&h[rank],mine));
    for(int r=0;r<2;++r) MPI_Bcast(&h[r],sizeof(h[r]),MPI_BYTE,r,MPI_COMM_WORLD);
    void* peer; CK(cudaIpcOpenMemHandle(&peer,h[1-rank],cudaIpcMemLazyEnablePeerAccess));
    double probe=0; CK(cudaMemcpy(&probe,(char*)peer+bytes-8,8,cudaMemcpyDeviceToHost)); // like the tag check
    if(variant==1||variant==3){ fillk<<<64,256,0,s2>>>(mine,n,1000.0*(rank+1)); CK(cudaStreamSynchronize(s2)); }
    if(variant==2){ std::vector<double> hb(n); for(size_t i=0;i<n;++i) hb[i]=1000.0*(rank+1)+i; CK(cudaMemcpy(mine,hb.data(),bytes,cudaMemcpyHostToDevice)); }
    if(variant==3){ CK(cudaDeviceSynchronize()); }
    double* dst; CK(cudaMalloc(&dst,bytes)); CK(cudaMemset(dst,0,bytes)); CK(cudaDeviceSynchronize());
    MPI_Barrier(MPI_COMM_WORLD);
    double t0=now(); CK(cudaMemcpyAsync(dst,peer,bytes,cudaMemcpyDeviceToDevice,s)); CK(cudaStreamSynchronize(s)); double t_copy=now()-t0;
    std::vector<double> hbuf(n); CK(cudaMemcpy(hbuf.data(),dst,bytes,cudaMemcpyDeviceToHost));
    size_t bad=0; for(size_t i=0;i<n;++i) if(hbuf[i]!=1000.0*(2-rank)+i) ++bad;
    const char* names[4]={"kernel fill BEFORE open","kernel fill AFTER open","H2D copy AFTER open","kernel fill AFTER open + device sync"};
    printf("[%d] %-40s copy %.3f ms bad=%zu first=%.1f\n",rank,names[variant],t_copy*1e3,bad,hbuf[0]);
    MPI_Barrier(MPI_COMM_WORLD);
    CK(cudaIpcCloseMemHandle(peer)); MPI_Barrier(MPI_COMM_WORLD); CK(cudaFree(mine)); CK(cudaFree(dst));
  }
  MPI_Finalize(); return 0; }
