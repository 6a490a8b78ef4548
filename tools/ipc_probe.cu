// ipc_probe.cu — CUDA IPC peer-pull sanity check under the MPI shim (2 ranks, 2 GPUs).
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
  size_t sizes[3]={ (size_t)300<<10, (size_t)2<<20, (size_t)64<<20 };
  for(int t=0;t<3;++t){
    size_t bytes=sizes[t], n=bytes/8; double *mine,*other_small; 
    CK(cudaMalloc(&other_small, 4096)); CK(cudaMalloc(&mine,bytes));
    fillk<<<64,256>>>(mine,n,1000.0*(rank+1)); CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h[2]; CK(cudaIpcGetMemHandle(&h[rank],mine));
    for(int r=0;r<2;++r) MPI_Bcast(&h[r],sizeof(h[r]),MPI_BYTE,r,MPI_COMM_WORLD);
    double t0=now(); void* peer; CK(cudaIpcOpenMemHandle(&peer,h[1-rank],cudaIpcMemLazyEnablePeerAccess)); double t_open=now()-t0;
    double* dst; CK(cudaMalloc(&dst,bytes)); CK(cudaMemset(dst,0,bytes));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s,cudaStreamNonBlocking));
    MPI_Barrier(MPI_COMM_WORLD);
    t0=now(); CK(cudaMemcpyAsync(dst,peer,bytes,cudaMemcpyDeviceToDevice,s)); CK(cudaStreamSynchronize(s)); double t_copy=now()-t0;
    std::vector<double> hbuf(n); CK(cudaMemcpy(hbuf.data(),dst,bytes,cudaMemcpyDeviceToHost));
    size_t bad=0; for(size_t i=0;i<n;++i) if(hbuf[i]!=1000.0*(2-rank)+i) ++bad;
    printf("[%d] size %zu KiB: open %.3f s, copy %.3f ms (%.1f GB/s), bad=%zu first=%.1f\n",rank,bytes>>10,t_open,t_copy*1e3,bytes/t_copy/1e9,bad,hbuf[0]);
    MPI_Barrier(MPI_COMM_WORLD);
    CK(cudaIpcCloseMemHandle(peer)); MPI_Barrier(MPI_COMM_WORLD); CK(cudaFree(mine)); CK(cudaFree(dst)); CK(cudaFree(other_small));
  }
  MPI_Finalize(); return 0; }
