// cuda_emu.cpp -- fiber scheduler behind tests/emu/cuda_emu.h (TEST-ONLY, see that header).
#include "cuda_emu.h"

#include <sys/mman.h>

namespace emu {
dim3 g_gridDim, g_blockDim, g_blockIdx, g_threadIdx;
char* g_dyn_smem = nullptr;

namespace {
struct Fiber {
  void* sp = nullptr;  // saved stack pointer
  char* stack = nullptr;
  bool done = false;
  dim3 tid;
};
constexpr size_t kStack = 256 * 1024;
std::vector<Fiber> g_fibers;
void* g_sched_sp = nullptr;
int g_cur = -1;
const std::function<void()>* g_body = nullptr;
int g_block_count = 0, g_block_gen = 0;
int g_warp_count[64], g_warp_gen[64];
float g_warp_slot[64][32];
std::vector<char> g_smem_buf;

// minimal x86-64 SysV context switch: save callee-saved regs on the current stack, swap rsp.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");

void yield_to_sched() {
  Fiber& f = g_fibers[g_cur];
  emu_switch(&f.sp, g_sched_sp);
}

void fiber_entry() {
  (*g_body)();
  g_fibers[g_cur].done = true;
  yield_to_sched();
  abort();  // never resumed
}

void prepare(Fiber& f) {
  if (!f.stack) {
    f.stack = static_cast<char*>(mmap(nullptr, kStack, PROT_READ | PROT_WRITE,
                                      MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
    if (f.stack == MAP_FAILED) { perror("mmap"); abort(); }
  }
  // stack layout for the first emu_switch into this fiber: 6 callee-saved slots + return address.
  uintptr_t top = reinterpret_cast<uintptr_t>(f.stack + kStack);
  top &= ~uintptr_t(15);
  void** sp = reinterpret_cast<void**>(top);
  *--sp = nullptr;                                   // fake return address of fiber_entry (alignment)
  *--sp = reinterpret_cast<void*>(&fiber_entry);     // `ret` target
  for (int i = 0; i < 6; ++i) *--sp = nullptr;       // rbp rbx r12..r15
  f.sp = sp;
  f.done = false;
}
}  // namespace

int lane() {
  int lin = g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z);
  return lin & 31;
}
static int linear_tid() { return g_threadIdx.x + g_blockDim.x * (g_threadIdx.y + g_blockDim.y * g_threadIdx.z); }

void sync_block() {
  int n = g_blockDim.x * g_blockDim.y * g_blockDim.z;
  int gen = g_block_gen;
  if (++g_block_count == n) { g_block_count = 0; ++g_block_gen; return; }
  while (g_block_gen == gen) yield_to_sched();
}

void sync_warp() {
  int n = g_blockDim.x * g_blockDim.y * g_blockDim.z;
  int w = linear_tid() >> 5;
  int expected = std::min(32, n - w * 32);
  int gen = g_warp_gen[w];
  if (++g_warp_count[w] == expected) { g_warp_count[w] = 0; ++g_warp_gen[w]; return; }
  while (g_warp_gen[w] == gen) yield_to_sched();
}

float shfl_f(float v, int src_lane) {
  int w = linear_tid() >> 5;
  g_warp_slot[w][lane()] = v;
  sync_warp();
  float r = g_warp_slot[w][src_lane & 31];
  sync_warp();
  return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  int n = block.x * block.y * block.z;
  if (n > 2048 || n <= 0) { fprintf(stderr, "emu: bad block size %d\n", n); abort(); }
  if ((int)g_fibers.size() < n) g_fibers.resize(n);
  if (g_smem_buf.size() < smem + 1024) g_smem_buf.resize(smem + 1024);
  g_dyn_smem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(g_smem_buf.data()) + 1023) & ~uintptr_t(1023));
  g_gridDim = grid;
  g_blockDim = block;
  g_body = &body;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_blockIdx = dim3(bx, by, bz);
        g_block_count = 0;
        memset(g_warp_count, 0, sizeof g_warp_count);
        int i = 0;
        for (unsigned tz = 0; tz < block.z; ++tz)
          for (unsigned ty = 0; ty < block.y; ++ty)
            for (unsigned tx = 0; tx < block.x; ++tx) {
              prepare(g_fibers[i]);
              g_fibers[i].tid = dim3(tx, ty, tz);
              ++i;
            }
        int remaining = n;
        long spins = 0;
        while (remaining > 0) {
          for (int f = 0; f < n; ++f) {
            if (g_fibers[f].done) continue;
            g_cur = f;
            g_threadIdx = g_fibers[f].tid;
            emu_switch(&g_sched_sp, g_fibers[f].sp);
            if (g_fibers[f].done) --remaining;
          }
          if (++spins > 100000000L) { fprintf(stderr, "emu: deadlock\n"); abort(); }
        }
      }
  g_body = nullptr;
}
}  // namespace emu
