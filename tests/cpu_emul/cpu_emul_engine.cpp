// TEST INFRASTRUCTURE - NOT PRODUCT CODE (see cuda_runtime.h in this directory).
// SIMT emulator: the CUDA threads of one block are fibers on one host thread, switched by a six-register context switch;
// __syncthreads / __syncwarp are counting barriers between the fibers that are still alive (a thread that has returned does
// not take part, as on sm_70+), __shfl_xor_sync exchanges through a per-warp slot array between two warp barriers. Kernels the
// preprocessor found free of barrier primitives run their threads as plain calls; a barrier reached in that mode aborts.
#include "cuda_runtime.h"

#include <sys/mman.h>

#include <vector>

#include <map>

namespace cpu_emul
{
  namespace
  {
    constexpr size_t kZone = 256;
    constexpr unsigned char kPattern = 0xA5;
    struct Live
    {
      std::map<char *, size_t> m;
      static void check(char *user, size_t n, const char *when)
      {
        for (size_t i = 0; i < kZone; ++i)
          if ((unsigned char)user[-(long)kZone + (long)i] != kPattern || (unsigned char)user[n + i] != kPattern)
            {
              std::fprintf(stderr, "cpu_emul: RED ZONE of a %zu-byte device buffer overwritten (%s the buffer, offset %zu; found at %s)\n", n,
                           (unsigned char)user[n + i] != kPattern ? "after" : "before", i, when);
              std::abort();
            }
      }
      ~Live()
      {
        for (auto &kv : m) check(kv.first, kv.second, "process exit");
      }
    };
    Live &live()
    {
      static Live l;
      return l;
    }
  } // namespace

  void *guarded_alloc(size_t n)
  {
    const size_t padded = (n + 63) / 64 * 64;
    char *raw = (char *)std::aligned_alloc(64, padded + 2 * kZone);
    if (!raw) return nullptr;
    std::memset(raw, kPattern, padded + 2 * kZone);
    char *user = raw + kZone;
    live().m[user] = n;
    return user;
  }

  void guarded_free(void *p)
  {
    if (!p) return;
    char *user = (char *)p;
    auto it = live().m.find(user);
    if (it == live().m.end())
      {
        std::fprintf(stderr, "cpu_emul: cudaFree of a pointer cudaMalloc did not return\n");
        std::abort();
      }
    Live::check(user, it->second, "cudaFree");
    live().m.erase(it);
    std::free(user - kZone);
  }

  Idx g_tid{0, 0, 0}, g_bid{0, 0, 0};
  dim3 g_block, g_grid;
  long long g_launches = 0, g_fiber_launches = 0;
} // namespace cpu_emul

extern "C" void cpu_emul_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cpu_emul_switch
.type cpu_emul_switch,@function
cpu_emul_switch:
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
.size cpu_emul_switch,.-cpu_emul_switch
)");

namespace cpu_emul
{
  namespace
  {
    enum State { READY, BLOCK_WAIT, WARP_WAIT, DONE };
    constexpr size_t kStack = 256 * 1024;
    constexpr int kMaxThreads = 1024;

    struct Fiber
    {
      void *sp;
      State state;
    };

    struct Engine
    {
      char *stacks = nullptr;
      Fiber fibers[kMaxThreads];
      void *main_sp = nullptr;
      int n = 0, cur = -1, alive = 0, block_waiting = 0;
      int warp_alive[kMaxThreads / 32], warp_waiting[kMaxThreads / 32];
      uint64_t slots[kMaxThreads];
      const std::function<void()> *body = nullptr;
      std::vector<char> smem;
      bool fiber_mode = false;
      const char *what = "";
    } E;

    void die(const char *msg)
    {
      std::fprintf(stderr, "cpu_emul: %s (kernel launch: %s)\n", msg, E.what);
      std::abort();
    }

    void release_block()
    {
      for (int i = 0; i < E.n; ++i)
        if (E.fibers[i].state == BLOCK_WAIT) E.fibers[i].state = READY;
      E.block_waiting = 0;
    }
    void release_warp(int w)
    {
      for (int i = 32 * w; i < 32 * w + 32 && i < E.n; ++i)
        if (E.fibers[i].state == WARP_WAIT) E.fibers[i].state = READY;
      E.warp_waiting[w] = 0;
    }

    void fiber_entry()
    {
      (*E.body)();
      const int i = E.cur, w = i / 32;
      E.fibers[i].state = DONE;
      --E.alive;
      --E.warp_alive[w];
      // a thread that returns completes the barriers the others are waiting at
      if (E.block_waiting > 0 && E.block_waiting == E.alive) release_block();
      if (E.warp_waiting[w] > 0 && E.warp_waiting[w] == E.warp_alive[w]) release_warp(w);
      void *dummy;
      cpu_emul_switch(&dummy, E.main_sp);
      die("a finished fiber was resumed");
    }

    void yield_to_scheduler() { cpu_emul_switch(&E.fibers[E.cur].sp, E.main_sp); }

    // IFEM_EMUL_SHUFFLE=<seed>: threads of a block (and the blocks of a launch) run in a pseudo-random order instead of
    // 0, 1, 2, ... - code that only works because a producer lane happens to run before its consumer (a missing barrier)
    // then gives different results. A sequential schedule still cannot show lost updates of unsynchronised read-modify-writes.
    uint64_t shuffle_seed()
    {
      static const uint64_t s = [] {
        const char *e = std::getenv("IFEM_EMUL_SHUFFLE");
        return e ? (uint64_t)std::strtoull(e, nullptr, 10) : 0ull;
      }();
      return s;
    }
    uint64_t g_rng = 0x9E3779B97F4A7C15ull;
    void shuffle(int *a, int n)
    {
      for (int i = n - 1; i > 0; --i)
        {
          g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull + shuffle_seed();
          const int j = (int)((g_rng >> 33) % (uint64_t)(i + 1));
          const int t = a[i];
          a[i] = a[j];
          a[j] = t;
        }
    }

    void run_block_fibers(unsigned n)
    {
      if (n > (unsigned)kMaxThreads) die("block larger than 1024 threads");
      if (!E.stacks)
        {
          E.stacks = (char *)mmap(nullptr, kStack * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
          if (E.stacks == (char *)MAP_FAILED) die("cannot map fiber stacks");
        }
      E.n = (int)n;
      E.alive = (int)n;
      E.block_waiting = 0;
      for (unsigned w = 0; w < (n + 31) / 32; ++w)
        {
          E.warp_alive[w] = (int)std::min(32u, n - 32 * w);
          E.warp_waiting[w] = 0;
        }
      for (unsigned i = 0; i < n; ++i)
        {
          char *top = E.stacks + kStack * (i + 1);
          void **sp = (void **)(top - 64);
          for (int k = 0; k < 6; ++k) sp[k] = nullptr;
          sp[6] = (void *)&fiber_entry;
          sp[7] = nullptr;
          E.fibers[i].sp = sp;
          E.fibers[i].state = READY;
        }
      std::vector<int> order(E.n);
      for (int i = 0; i < E.n; ++i) order[i] = i;
      while (E.alive > 0)
        {
          bool progressed = false;
          if (shuffle_seed()) shuffle(order.data(), E.n); // a different thread order in every pass (IFEM_EMUL_SHUFFLE)
          for (int k = 0; k < E.n; ++k)
            {
              const int i = order[k];
              if (E.fibers[i].state != READY) continue;
              progressed = true;
              E.cur = i;
              g_tid = Idx{(unsigned)i, 0, 0};
              cpu_emul_switch(&E.main_sp, E.fibers[i].sp);
            }
          if (!progressed) die("deadlock: every live thread waits at a barrier that cannot complete");
        }
      E.cur = -1;
    }
  } // namespace

  void *dyn_smem() { return E.smem.data(); }

  void sync_threads()
  {
    if (!E.fiber_mode) die("__syncthreads reached in a kernel classified as barrier-free");
    if (E.block_waiting + 1 == E.alive) // the last live thread to arrive releases the others and goes on
      {
        release_block();
        return;
      }
    ++E.block_waiting;
    E.fibers[E.cur].state = BLOCK_WAIT;
    yield_to_scheduler();
  }

  void sync_warp()
  {
    if (!E.fiber_mode) die("__syncwarp / warp shuffle reached in a kernel classified as barrier-free");
    const int w = E.cur / 32;
    if (E.warp_waiting[w] + 1 == E.warp_alive[w])
      {
        release_warp(w);
        return;
      }
    ++E.warp_waiting[w];
    E.fibers[E.cur].state = WARP_WAIT;
    yield_to_scheduler();
  }

  uint64_t shfl_xor_bits(uint64_t bits, int lane_mask)
  {
    const int me = E.cur;
    E.slots[me] = bits;
    sync_warp();
    const int partner = (me & ~31) | ((me ^ lane_mask) & 31);
    const uint64_t r = partner < E.n ? E.slots[partner] : bits;
    sync_warp();
    return r;
  }

  void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, bool uses_barriers, const char *what, const std::function<void()> &body)
  {
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1) die("only 1-D launches are emulated");
    if (E.body) die("nested kernel launch");
    ++g_launches;
    if (uses_barriers) ++g_fiber_launches;
    g_grid = grid;
    g_block = block;
    E.what = what;
    E.body = &body;
    E.fiber_mode = uses_barriers;
    if (E.smem.size() < dyn_smem_bytes + 64) E.smem.resize(dyn_smem_bytes + 64);
    std::vector<int> blocks(grid.x), threads(block.x);
    for (unsigned b = 0; b < grid.x; ++b) blocks[b] = (int)b;
    for (unsigned t = 0; t < block.x; ++t) threads[t] = (int)t;
    if (shuffle_seed()) shuffle(blocks.data(), (int)grid.x);
    for (unsigned bi = 0; bi < grid.x; ++bi)
      {
        g_bid = Idx{(unsigned)blocks[bi], 0, 0};
        if (uses_barriers)
          run_block_fibers(block.x);
        else
          {
            if (shuffle_seed()) shuffle(threads.data(), (int)block.x);
            for (unsigned t = 0; t < block.x; ++t)
              {
                g_tid = Idx{(unsigned)threads[t], 0, 0};
                body();
              }
          }
      }
    E.body = nullptr;
    E.fiber_mode = false;
  }
} // namespace cpu_emul
