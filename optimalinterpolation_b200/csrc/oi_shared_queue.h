// A double-ended work queue shared by the GPU processes of ONE box (one process per GPU, torchrun): a single 64-bit
// word in POSIX shared memory holding {run generation, front cursor, back cursor} over the cost-sorted list of pending
// cells that every rank builds identically (all ranks hold all observations and all cells of the step, as every MPI rank of
// the reference does, GPR_CS2S3.py:201-246).  A rank's bulk groups claim the largest unclaimed cell from the front, its
// small-first groups the smallest from the back (oi_api.cu: LockstepRun), so the static split of the reference
// (`split(container, count)`, :18-23) becomes self-scheduling on the same cost-sorted list: every GPU stays busy until the
// list is empty, whatever the cells' evaluation counts turn out to be.  No collective, no NCCL: the only collective of the
// path stays the final gather of the result rows.  Per-cell results do not depend on which rank computes a cell.
//
// Word layout: [63:48] generation (run counter mod 2^16) | [47:24] front | [23:0] back    (< 2^24 pending cells)
// Every oi_run of a step uses the same generation on all ranks; the first rank that arrives (CAS) resets the cursors to
// {0, n_pending}.  A rank can only move on to generation g+1 after the list of generation g was exhausted, so a late rank
// that still holds generation g sees a NEWER word and treats the list as empty.
//
// Pure C++11 + POSIX (also compiled into the host test driver, tests/cg_host.cpp).
#pragma once
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

struct OiSharedQueue {
    std::atomic<uint64_t>* word = nullptr;
    int fd = -1;
    char name[128] = {0};
    uint32_t gen = 0;          // generation of the current run (this process)
    uint32_t n = 0;            // pending cells of the current run

    static uint64_t pack(uint32_t g, uint32_t f, uint32_t b) { return ((uint64_t)(g & 0xffffu) << 48) | ((uint64_t)(f & 0xffffffu) << 24) | (uint64_t)(b & 0xffffffu); }
    static uint32_t gen_of(uint64_t v) { return (uint32_t)(v >> 48) & 0xffffu; }
    static uint32_t front_of(uint64_t v) { return (uint32_t)(v >> 24) & 0xffffffu; }
    static uint32_t back_of(uint64_t v) { return (uint32_t)v & 0xffffffu; }

    bool attached() const { return word != nullptr; }

    // returns 0 on success, errno otherwise; a new segment starts zero-filled (generation 0, empty)
    int attach(const char* shm_name) {
        detach();
        int f = shm_open(shm_name, O_CREAT | O_RDWR, 0600);
        if (f < 0) return errno ? errno : -1;
        if (ftruncate(f, 64) != 0) { int e = errno; close(f); return e ? e : -1; }
        void* p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, f, 0);
        if (p == MAP_FAILED) { int e = errno; close(f); return e ? e : -1; }
        static_assert(sizeof(std::atomic<uint64_t>) == 8, "lock-free 64-bit atomics expected");
        word = reinterpret_cast<std::atomic<uint64_t>*>(p);
        fd = f;
        std::strncpy(name, shm_name, sizeof(name) - 1);
        gen = 0;       // every rank attaches to a FRESH segment (unique name) before the first run of the job
        return 0;
    }
    void detach() {
        if (word) munmap((void*)word, 64);
        if (fd >= 0) close(fd);
        word = nullptr; fd = -1;
    }
    static void unlink_name(const char* shm_name) { shm_unlink(shm_name); }

    // Start a run over n_pending cells: all ranks call it once per run, in the same order of runs.
    void begin_run(uint32_t n_pending) {
        gen = (gen + 1) & 0xffffu;
        n = n_pending;
        uint64_t v = word->load(std::memory_order_acquire);
        while (older(gen_of(v), gen)) {
            if (word->compare_exchange_weak(v, pack(gen, 0, n_pending), std::memory_order_acq_rel)) break;
        }
    }
    // a is an older generation than b (mod 2^16)
    static bool older(uint32_t a, uint32_t b) { uint32_t d = (b - a) & 0xffffu; return d != 0 && d < 0x8000u; }

    // Index of the largest unclaimed cell, or -1 when the list is exhausted (or belongs to a newer run).  Does not claim.
    long peek_front() const {
        uint64_t v = word->load(std::memory_order_acquire);
        if (gen_of(v) != gen || front_of(v) >= back_of(v)) return -1;
        return (long)front_of(v);
    }
    long peek_back() const {
        uint64_t v = word->load(std::memory_order_acquire);
        if (gen_of(v) != gen || front_of(v) >= back_of(v)) return -1;
        return (long)back_of(v) - 1;
    }
    // Claim exactly the index a peek returned; false = somebody else moved the cursor first (peek again).
    bool claim_front(long idx) {
        uint64_t v = word->load(std::memory_order_acquire);
        if (gen_of(v) != gen || (long)front_of(v) != idx || front_of(v) >= back_of(v)) return false;
        return word->compare_exchange_strong(v, pack(gen, front_of(v) + 1, back_of(v)), std::memory_order_acq_rel);
    }
    bool claim_back(long idx) {
        uint64_t v = word->load(std::memory_order_acquire);
        if (gen_of(v) != gen || (long)back_of(v) - 1 != idx || front_of(v) >= back_of(v)) return false;
        return word->compare_exchange_strong(v, pack(gen, front_of(v), back_of(v) - 1), std::memory_order_acq_rel);
    }
    bool exhausted() const { return peek_front() < 0; }
};
