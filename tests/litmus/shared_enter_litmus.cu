// Litmus test of the ordered-shared-section protocol of hived_core.h (sharedEnter / the release at the end of an
// event), with the library's own primitives (hived_prims.h) — test infrastructure, built by tests/test_gpu_parity.py.
//
// The protocol is message passing between CTAs through global memory:
//   writer (an event that touched the cluster-wide state):   plain stores of the warp's lanes to the state;
//                                                            hv_fence(); lane 0: hv_st_volatile(progress, next)
//   reader (sharedEnter of a later event):                   spin: hv_ld_volatile(progress) > myEvent;
//                                                            hv_fence(); plain loads of the state
// The reader's loads may hit L1 lines it filled BEFORE the writer wrote (the leader warp keeps the state in L1 on
// purpose): the test makes that the common case by reading every word right before it waits.  It reports how many
// stale words the reader saw (a) with the protocol as the library runs it and (b) with the reader's fence left out
// (the control: shows that the test can see the failure at all — it may legitimately pass on a given part).
//
//   usage: shared_enter_litmus [rounds] [words]      exit code 0 iff (a) saw no stale word
#include <cstdio>
#include <cstdlib>

#include "../../hivedscheduler_b200/csrc/hived_prims.h"

using namespace hived;

__global__ void litmus(int* state, int words, int* progress, int rounds, int readerFence, unsigned long long* stale,
                       unsigned long long* checked) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x >= 32) return;
  // CTA 0 writes in odd rounds, CTA 1 in even rounds: each is reader and writer in turn, like two VCs' CTAs
  const int me = blockIdx.x;
  unsigned long long bad = 0, seen = 0;
  for (int r = 1; r <= rounds; r++) {
    const bool writer = (r & 1) == me;
    if (writer) {
      for (int i = lane; i < words; i += 32) state[i] = r;     // the shared section
      __syncwarp();
      hv_fence();                                              // release (run(): `if (sharedHeld) hv_fence()`)
      if (lane == 0) hv_st_volatile(progress + me, r);
      __syncwarp();
    } else {
      // pull the (old) state into L1, as the leader warp's ordinary work does
      int sum = 0;
      for (int i = lane; i < words; i += 32) sum += state[i];
      if (sum == 0x7fffffff) state[0] = 0;                     // (keeps the loads alive)
      while (true) {                                           // sharedEnterSlow
        int v = lane == 0 ? hv_ld_volatile(progress + (me ^ 1)) : 0;
        v = __shfl_sync(0xffffffffu, v, 0);
        if (v >= r) break;
      }
      if (readerFence) hv_fence();                             // acquire
      for (int i = lane; i < words; i += 32) { seen++; if (state[i] != r) bad++; }
      __syncwarp();
      // tell the writer that the round was read (so that it may overwrite): same protocol the other way round
      hv_fence();
      if (lane == 0) hv_st_volatile(progress + me, r);
      __syncwarp();
    }
    // both wait until the other has finished the round
    while (true) {
      int v = lane == 0 ? hv_ld_volatile(progress + (me ^ 1)) : 0;
      v = __shfl_sync(0xffffffffu, v, 0);
      if (v >= r) break;
    }
    hv_fence();
  }
  atomicAdd(stale, bad);
  atomicAdd(checked, seen);
}

static int run(int rounds, int words, int readerFence, unsigned long long* staleOut) {
  int *state, *progress;
  unsigned long long *cnt, h[2] = {0, 0};
  cudaMalloc(&state, (size_t)words * 4);
  cudaMalloc(&progress, 64);
  cudaMalloc(&cnt, 16);
  cudaMemset(state, 0, (size_t)words * 4);
  cudaMemset(progress, 0, 64);
  cudaMemset(cnt, 0, 16);
  void* args[] = {&state, &words, &progress, &rounds, &readerFence, nullptr, nullptr};
  unsigned long long* a5 = cnt;
  unsigned long long* a6 = cnt + 1;
  args[5] = &a5;
  args[6] = &a6;
  // the two CTAs wait for each other: cooperative launch, as the library does for a VC-parallel batch
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)litmus, dim3(2), dim3(64), args, 0, 0);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "litmus: %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(h, cnt, 16, cudaMemcpyDeviceToHost);
  cudaFree(state); cudaFree(progress); cudaFree(cnt);
  *staleOut = h[0];
  printf("reader fence %s: rounds %d, words %d, words checked %llu, stale words %llu\n", readerFence ? "on " : "off", rounds, words, h[1], h[0]);
  return 0;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 20000, words = argc > 2 ? atoi(argv[2]) : 4096;
  unsigned long long staleOn = 0, staleOff = 0;
  if (run(rounds, words, 1, &staleOn)) return 2;
  if (run(rounds, words, 0, &staleOff)) return 2;
  printf("shared_enter_litmus: %s (control without the acquire fence saw %llu stale words)\n", staleOn == 0 ? "ok" : "FAILED", staleOff);
  return staleOn == 0 ? 0 : 1;
}
