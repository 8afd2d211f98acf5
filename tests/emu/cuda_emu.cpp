// cuda_emu.cpp — scheduler of the fiber SIMT emulator (see cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"

namespace zrk_emu {

Block *g_block = nullptr;
dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

static void fiber_entry() {
    Block *b = g_block;
    (*b->body)();
    b->done[b->cur] = 1;
    b->n_done++;
    // a thread that exits releases anyone waiting on __syncthreads for it
    Barrier &bar = b->named[0];
    if (bar.count > 0 && bar.count >= b->n_threads - b->n_done) {
        bar.count = 0;
        bar.gen++;
    }
    swapcontext(&b->ctx[b->cur], &b->main_ctx);
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    const size_t STACK = 256 * 1024;
    const int n = (int)block.x;
    if (n % 32) {
        fprintf(stderr, "zrk_emu: block size must be a multiple of 32\n");
        abort();
    }
    Block b;
    b.n_threads = n;
    b.ctx.resize(n);
    b.done.assign(n, 0);
    b.stacks.resize(n);
    b.warps.resize(n / 32);
    b.body = &body;
    for (int i = 0; i < n; i++) b.stacks[i] = (char *)malloc(STACK);
    void *sm = nullptr;
    if (posix_memalign(&sm, 128, smem_bytes ? smem_bytes : 16)) abort();
    b.smem = (unsigned char *)sm;
    g_blockDim = block;
    g_gridDim = grid;
    Block *saved = g_block;
    g_block = &b;
    for (unsigned bx = 0; bx < grid.x; bx++) {
        g_blockIdx = dim3(bx, 0, 0);
        b.n_done = 0;
        for (int i = 0; i < 17; i++) b.named[i] = Barrier();
        for (auto &w : b.warps) w = Warp();
        memset(b.smem, 0xA5, smem_bytes);  // shared memory is uninitialised on the GPU
        for (int i = 0; i < n; i++) {
            b.done[i] = 0;
            getcontext(&b.ctx[i]);
            b.ctx[i].uc_stack.ss_sp = b.stacks[i];
            b.ctx[i].uc_stack.ss_size = STACK;
            b.ctx[i].uc_link = &b.main_ctx;
            makecontext(&b.ctx[i], fiber_entry, 0);
        }
        long spins = 0;
        while (b.n_done < n) {
            for (int i = 0; i < n; i++) {
                if (b.done[i]) continue;
                b.cur = i;
                g_threadIdx = dim3((unsigned)i, 0, 0);
                swapcontext(&b.main_ctx, &b.ctx[i]);
            }
            if (++spins > 200000000L) {
                fprintf(stderr, "zrk_emu: deadlock suspected in block %u\n", bx);
                abort();
            }
        }
    }
    g_block = saved;
    for (int i = 0; i < n; i++) free(b.stacks[i]);
    free(b.smem);
}

}  // namespace zrk_emu
