// Where does the per-call latency of the synchronous host API go? Floors of this box, measured from the host:
//   (1) empty kernel launch + cudaStreamSynchronize
//   (2) a kernel that reads 8 KB from and writes 8 KB to PINNED host memory (what the zero-copy path of
//       b2f_cfft_exec_host does for one 1024-point transform), + cudaStreamSynchronize
//   (3) the same, completion signalled by a flag the kernel writes to pinned memory and the host polls
//   (4) the same with H2D / D2H cudaMemcpyAsync around a device-memory kernel
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/launch_floor.cu -o tools/launch_floor.bin
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__global__ void empty_kernel() {}
__global__ void copy_kernel(const float4 *in, float4 *out, int n4) {
  for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = in[i];
}
__global__ void copy_flag_kernel(const float4 *in, float4 *out, int n4, volatile unsigned *flag, unsigned seq) {
  for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = in[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *flag = seq;
}
// (6) launched AHEAD of the call: spins on a host flag until the host says go
__global__ void gated_copy_kernel(const float4 *in, float4 *out, int n4, volatile unsigned *go, volatile unsigned *done, unsigned seq) {
  if (threadIdx.x == 0)
    while (*go < seq) {
    }
  __syncthreads();
  for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = in[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *done = seq;
}
// (7) resident: one launch serves every call until told to stop (seq 0xffffffff)
__global__ void resident_copy_kernel(const float4 *in, float4 *out, int n4, volatile unsigned *go, volatile unsigned *done) {
  __shared__ unsigned cur;
  unsigned seq = 1;
  for (;;) {
    if (threadIdx.x == 0) {
      unsigned g;
      while ((g = *go) < seq) {
      }
      cur = g;
    }
    __syncthreads();
    if (cur == 0xffffffffu) return;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = in[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *done = seq;
    seq++;
  }
}
template <class F>
static double time_us(F f, int n = 2000) {
  for (int i = 0; i < 100; i++) f(i);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++) f(100 + i);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::micro>(t1 - t0).count() / n;
}
int main() {
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int bytes = 8192, n4 = bytes / 16;
  float4 *hin, *hout, *din, *dout;
  unsigned *flag;
  cudaMallocHost(&hin, bytes);
  cudaMallocHost(&hout, bytes);
  cudaMallocHost(&flag, 64);
  cudaMalloc(&din, bytes);
  cudaMalloc(&dout, bytes);
  memset(hin, 1, bytes);
  *flag = 0;
  printf("empty launch + sync                 %6.2f us\n", time_us([&](int) {
           empty_kernel<<<1, 32, 0, st>>>();
           cudaStreamSynchronize(st);
         }));
  printf("pinned in/out kernel + sync         %6.2f us\n", time_us([&](int) {
           copy_kernel<<<1, 256, 0, st>>>(hin, hout, n4);
           cudaStreamSynchronize(st);
         }));
  printf("pinned in/out kernel + polled flag  %6.2f us\n", time_us([&](int i) {
           copy_flag_kernel<<<1, 256, 0, st>>>(hin, hout, n4, flag, (unsigned)i + 1);
           while (*(volatile unsigned *)flag != (unsigned)i + 1) {
           }
         }));
  printf("H2D + device kernel + D2H + sync    %6.2f us\n", time_us([&](int) {
           cudaMemcpyAsync(din, hin, bytes, cudaMemcpyHostToDevice, st);
           copy_kernel<<<1, 256, 0, st>>>(din, dout, n4);
           cudaMemcpyAsync(hout, dout, bytes, cudaMemcpyDeviceToHost, st);
           cudaStreamSynchronize(st);
         }));
  printf("device kernel + sync                %6.2f us\n", time_us([&](int) {
           copy_kernel<<<1, 256, 0, st>>>(din, dout, n4);
           cudaStreamSynchronize(st);
         }));
  // (5) stream memory operations: [wait go >= k] -> kernel -> [write done = k] enqueued BEFORE call k; call k is a flag
  // store, the enqueue of call k + 1's triple (overlapping the GPU's work) and a spin on the done flag
  typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
  WaitFn waitv = nullptr, writev = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuStreamWaitValue32", (void **)&waitv, cudaEnableDefault, &q);
  cudaGetDriverEntryPoint("cuStreamWriteValue32", (void **)&writev, cudaEnableDefault, &q);
  unsigned *go, *done;
  cudaHostAlloc(&go, 64, cudaHostAllocMapped);
  cudaHostAlloc(&done, 64, cudaHostAllocMapped);
  *go = 0;
  *done = 0;
  if (waitv && writev) {
    unsigned k = 1;
    auto arm = [&](unsigned seq) {
      waitv((CUstream)st, (CUdeviceptr)go, seq, CU_STREAM_WAIT_VALUE_GEQ);
      copy_kernel<<<1, 256, 0, st>>>(hin, hout, n4);
      writev((CUstream)st, (CUdeviceptr)done, seq, 0);
    };
    arm(k);
    printf("pre-armed memops, re-arm overlapped %6.2f us\n", time_us([&](int) {
             *(volatile unsigned *)go = k;
             arm(k + 1);
             while (*(volatile unsigned *)done != k) {
             }
             k++;
           }));
    *(volatile unsigned *)go = k;  // release the armed one
    cudaStreamSynchronize(st);
    // re-arm after completion (what a call that cannot overlap would pay)
    *go = 0, *done = 0, k = 1;
    arm(k);
    printf("pre-armed memops, re-arm afterwards %6.2f us\n", time_us([&](int) {
             *(volatile unsigned *)go = k;
             while (*(volatile unsigned *)done != k) {
             }
             arm(k + 1);
             k++;
           }));
    *(volatile unsigned *)go = k;
    cudaStreamSynchronize(st);
  }
  {
    *go = 0, *done = 0;
    unsigned k = 1;
    gated_copy_kernel<<<1, 256, 0, st>>>(hin, hout, n4, go, done, k);
    printf("pre-launched gated kernel           %6.2f us\n", time_us([&](int) {
             *(volatile unsigned *)go = k;
             gated_copy_kernel<<<1, 256, 0, st>>>(hin, hout, n4, go, done, k + 1);
             while (*(volatile unsigned *)done != k) {
             }
             k++;
           }));
    *(volatile unsigned *)go = k;
    cudaStreamSynchronize(st);
  }
  {
    *go = 0, *done = 0;
    unsigned k = 1;
    resident_copy_kernel<<<1, 256, 0, st>>>(hin, hout, n4, go, done);
    printf("resident kernel (mailbox)           %6.2f us\n", time_us([&](int) {
             *(volatile unsigned *)go = k;
             while (*(volatile unsigned *)done != k) {
             }
             k++;
           }));
    *(volatile unsigned *)go = 0xffffffffu;
    cudaStreamSynchronize(st);
  }
  return 0;
}
