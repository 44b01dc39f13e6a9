// ct_tiles.hpp -- the tile dispenser behind dynamic tile stealing (north_star: "dynamic tile stealing via an
// atomic counter").  Replaces the reference's static split yStep = H / numberOfThreads (raythread.cpp:576).
// Process-local std::atomic for the GPUs one process drives; a POSIX shared-memory int32 when the job runs
// one process per GPU (torchrun), so that all ranks pull from the same counter.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace cth {

struct TileCounter {               // hands out tile numbers 0,1,2,... to whoever asks first
    std::atomic<int32_t> local{0};
    int32_t *shared = nullptr;     // mmap'd, shared by all processes of the job
    std::string shm_name;
    int fd = -1;

    void open_shared(const std::string &name) {
        shm_name = name[0] == '/' ? name : "/" + name;
        fd = shm_open(shm_name.c_str(), O_CREAT | O_RDWR, 0600);
        if (fd < 0) throw std::runtime_error("shm_open(" + shm_name + ") failed");
        if (ftruncate(fd, 64) != 0) throw std::runtime_error("ftruncate on shared tile counter failed");
        void *p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (p == MAP_FAILED) throw std::runtime_error("mmap of shared tile counter failed");
        shared = static_cast<int32_t *>(p);
    }
    int32_t next() { return shared ? __atomic_fetch_add(shared, 1, __ATOMIC_RELAXED) : local.fetch_add(1, std::memory_order_relaxed); }
    void reset() { if (shared) __atomic_store_n(shared, 0, __ATOMIC_SEQ_CST); else local.store(0); }
    ~TileCounter() {
        if (shared) munmap(shared, 64);
        if (fd >= 0) close(fd);
    }
};


}  // namespace cth
