// TEST INFRASTRUCTURE ONLY — a functional SIMT emulator for the device program (csrc/hived_core.h).
//
// The 1-lane emulation (hived_emu.cpp) checks the scheduling LOGIC; it cannot see the warp-level
// code (ballots, match, shuffles, reductions, lane-partitioned loops) nor the CTA barrier protocol
// between the leader warp and the worker warps, nor the ordered shared sections between CTAs.
// This runtime executes the same source with the real geometry — 32 lanes per warp, several
// warps per CTA, several CTAs per launch — on ONE host thread: every CUDA thread is a fiber
// (hand-rolled x86-64 context switch), a warp collective / CTA barrier blocks the fiber until all
// its peers have arrived, and a volatile load (spin-wait on another CTA's progress word) yields.
// Any interleaving it produces is one the CUDA memory model allows (threads between two
// synchronisation points run in arbitrary order), so a divergence from the oracle here is a real
// bug of the SIMT code; it is not a proof of the absence of races (the interleaving is fixed).
#pragma once
#include <execinfo.h>
#include <sys/mman.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace simt {

constexpr int WARP = 32;

extern "C" void simt_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size simt_switch,.-simt_switch
)");

struct WarpState {
  int vals[WARP];
  int res[WARP];
  int arrived = 0;
  unsigned gen = 0;
};
struct CtaState {
  int arrived = 0;
  unsigned gen = 0;
  int nThreads = 0;
  int finished = 0;
  int arrivedN = 0;      // named barrier over the first n threads (bar.sync 1, n)
  unsigned genN = 0;
};

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  size_t stackBytes = 0;
  int cta = 0, tid = 0, warp = 0, lane = 0, nth = 0;
  bool done = false;
  bool blocked = false;
  int waitKind = 0;  // 1 warp collective, 2 CTA barrier (diagnostics)
  void* bt[12];
  long long nBar = 0, nCol = 0;
  int nbt = 0;
  WarpState* w = nullptr;
  CtaState* c = nullptr;
};

struct Runtime {
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  std::vector<CtaState> ctas;
  void* mainSp = nullptr;
  Fiber* cur = nullptr;
  void (*entry)(void*) = nullptr;
  void* entryArg = nullptr;
  long long switches = 0;
};
inline Runtime& rt() { static Runtime r; return r; }
inline bool tracing() { static int t = getenv("SIMT_TRACE") ? 1 : 0; return t != 0; }

inline void yield_to_scheduler() {
  Runtime& r = rt();
  Fiber* f = r.cur;
  r.switches++;
  simt_switch(&f->sp, r.mainSp);
}

[[noreturn]] inline void fiber_main() {
  Runtime& r = rt();
  r.entry(r.entryArg);
  Fiber* f = r.cur;
  f->done = true;
  f->c->finished++;
  // a thread that left the kernel no longer takes part in barriers (CUDA: exited threads count as arrived)
  yield_to_scheduler();
  fprintf(stderr, "simt: resumed a finished fiber\n");
  abort();
}
extern "C" void simt_trampoline();
asm(R"(
.text
.globl simt_trampoline
.type simt_trampoline,@function
simt_trampoline:
  andq $-16, %rsp
  call simt_fiber_main_c
  hlt
.size simt_trampoline,.-simt_trampoline
)");
extern "C" void simt_fiber_main_c() { fiber_main(); }

// ---- what the device program sees
inline int lane() { return rt().cur->lane; }
inline int tid() { return rt().cur->tid; }
inline int nth() { return rt().cur->nth; }
inline int warp() { return rt().cur->warp; }
inline int cta() { return rt().cur->cta; }

// all 32 lanes of the calling warp deposit `v`; the last one to arrive runs `fin(vals, res)`
template <typename Fin>
inline int warp_collective(int v, Fin fin) {
  Fiber* f = rt().cur;
  WarpState& w = *f->w;
  f->nCol++;
  w.vals[f->lane] = v;
  if (++w.arrived == WARP) {
    fin(w.vals, w.res);
    w.arrived = 0;
    w.gen++;
    // wake the peers
    Runtime& r = rt();
    Fiber* base = f - f->lane;
    for (int i = 0; i < WARP; i++) base[i].blocked = false;
    (void)r;
  } else {
    unsigned g = w.gen;
    f->blocked = true; f->waitKind = 1;
    if (tracing()) f->nbt = backtrace(f->bt, 12);
    while (w.gen == g) yield_to_scheduler();
  }
  return w.res[f->lane];
}
inline void cta_barrier() {
  Fiber* f = rt().cur;
  CtaState& c = *f->c;
  f->nBar++;
  if (++c.arrived >= c.nThreads - c.finished) {
    c.arrived = 0;
    c.gen++;
    Fiber* base = f - f->tid;
    for (int i = 0; i < c.nThreads; i++) base[i].blocked = false;
  } else {
    unsigned g = c.gen;
    f->blocked = true; f->waitKind = 2;
    if (tracing()) f->nbt = backtrace(f->bt, 12);
    while (c.gen == g) yield_to_scheduler();
  }
}

// bar.sync 1, n: the first n threads of the CTA (none of them has left the kernel while the others wait here)
inline void cta_barrier_n(int n) {
  Fiber* f = rt().cur;
  CtaState& c = *f->c;
  f->nBar++;
  if (f->tid >= n) { fprintf(stderr, "simt: thread %d outside the named barrier of %d threads\n", f->tid, n); abort(); }
  if (++c.arrivedN >= n) {
    c.arrivedN = 0;
    c.genN++;
    Fiber* base = f - f->tid;
    for (int i = 0; i < n; i++) base[i].blocked = false;
  } else {
    unsigned g = c.genN;
    f->blocked = true; f->waitKind = 2;
    if (tracing()) f->nbt = backtrace(f->bt, 12);
    while (c.genN == g) yield_to_scheduler();
  }
}

// ---- launch: nCta CTAs of nThreads threads, every thread runs entry(arg)
inline void launch(int nCta, int nThreads, void (*entry)(void*), void* arg, size_t stackBytes = 1u << 20) {
  Runtime& r = rt();
  if (nThreads % WARP) { fprintf(stderr, "simt: block size must be a multiple of 32\n"); abort(); }
  const int total = nCta * nThreads;
  r.fibers.assign(total, Fiber());
  r.warps.assign(total / WARP, WarpState());
  r.ctas.assign(nCta, CtaState());
  r.entry = entry;
  r.entryArg = arg;
  for (int c = 0; c < nCta; c++) r.ctas[c].nThreads = nThreads;
  for (int i = 0; i < total; i++) {
    Fiber& f = r.fibers[i];
    f.cta = i / nThreads; f.tid = i % nThreads; f.warp = f.tid / WARP; f.lane = f.tid % WARP; f.nth = nThreads;
    f.w = &r.warps[i / WARP];
    f.c = &r.ctas[f.cta];
    f.stackBytes = stackBytes;
    f.stack = (char*)mmap(nullptr, stackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (f.stack == MAP_FAILED) { fprintf(stderr, "simt: mmap failed\n"); abort(); }
    void** sp = (void**)(f.stack + stackBytes - 64);
    *--sp = (void*)simt_trampoline;  // return address of the first switch
    for (int k = 0; k < 6; k++) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = sp;
  }
  int alive = total;
  long long idle = 0;
  while (alive > 0) {
    bool ran = false;
    for (int i = 0; i < total; i++) {
      Fiber& f = r.fibers[i];
      if (f.done || f.blocked) continue;
      r.cur = &f;
      simt_switch(&r.mainSp, f.sp);
      ran = true;
      if (f.done) {
        alive--;
        // threads waiting on a CTA barrier that this thread will never reach
        CtaState& c = *f.c;
        if (c.arrived > 0 && c.arrived >= c.nThreads - c.finished) {
          c.arrived = 0; c.gen++;
          Fiber* base = &f - f.tid;
          for (int k = 0; k < c.nThreads; k++) base[k].blocked = false;
        }
      }
    }
    if (!ran) {
      if (++idle > 4) {
        fprintf(stderr, "simt: deadlock — every live thread is blocked (divergent collective or barrier?)\n");
        for (int i = 0; i < total; i++) {
          Fiber& f = r.fibers[i];
          fprintf(stderr, "%s%d:%d%c%lld/%lld", i % 8 ? " " : "\n", f.cta, f.tid, f.done ? 'D' : (f.waitKind == 1 ? 'w' : (f.waitKind == 2 ? 'B' : '?')), f.nBar, f.nCol);
        }
        fprintf(stderr, "\n");
        if (tracing())
          for (int i = 0; i < total && i < 64; i += 1) {
            Fiber& f = r.fibers[i];
            if (f.done || (i > 2 && i != 32 && i != 33)) continue;
            fprintf(stderr, "-- thread %d:%d barriers %lld collectives %lld (cta arrived %d)\n", f.cta, f.tid, f.nBar, f.nCol, f.c->arrived);
            backtrace_symbols_fd(f.bt, f.nbt, 2);
          }
        abort();
      }
    } else {
      idle = 0;
    }
  }
  for (auto& f : r.fibers) munmap(f.stack, f.stackBytes);
  r.fibers.clear();
  r.cur = nullptr;
}

}  // namespace simt
