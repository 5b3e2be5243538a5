// numa.cc — NUMA placement of the host side of a GPU's batch buffers, without libnuma.
//
// The reference's Evaluator can allocate its four batch arrays with numa_alloc_onnode and pin its thread to that
// node's CPUs (reference src/evaluate/evaluator.cc:39-83,127-136, NUMA_ENABLED builds).  On an 8-GPU box that is what
// keeps every GPU's D2H traffic - 8,756 B of dense logits per sample through the Infer contract, ~22 GB/s per GPU at
// 2.6 M evals/s - on its own socket's memory controllers instead of crossing the socket interconnect.  Here the node is
// not an argument: it is the GPU's own (sysfs numa_node of its PCI function), so one call places a buffer next to the
// GPU that will read and write it.
#include <cuda_runtime.h>
#include <ctype.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "nsb_internal.h"

namespace nsb {

// NUMA node of a CUDA device (-1: unknown or a single-node machine reporting -1)
int gpu_numa_node(int gpu) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, gpu) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char* p = bus; *p; ++p) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// CPUs of a node from /sys/devices/system/node/nodeN/cpulist ("0-31,64-95")
static bool node_cpus(int node, cpu_set_t* set) {
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char buf[4096] = "";
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* p = buf; *p;) {
        if (!isdigit((unsigned char)*p)) { ++p; continue; }
        long a = strtol(p, &p, 10), b = a;
        if (*p == '-') b = strtol(p + 1, &p, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++count; }
    }
    return count > 0;
}

int numa_bind_thread_to_gpu(int gpu) {
    const int node = gpu_numa_node(gpu);
    cpu_set_t set;
    if (node < 0 || !node_cpus(node, &set)) return 1;          // nothing to do on this machine
    // keep only CPUs this process may use at all (containers, taskset)
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0) {
        cpu_set_t both;
        CPU_AND(&both, &set, &allowed);
        if (CPU_COUNT(&both) == 0) return 1;
        set = both;
    }
    return sched_setaffinity(0, sizeof set, &set) == 0 ? 0 : 1;  // evaluator.cc:64-77
}

// Page-aligned anonymous memory placed on the GPU's node: mbind(MPOL_PREFERRED) when the kernel lets us, otherwise
// first touch from a thread temporarily pinned to the node's CPUs.  Returns nullptr on failure; *node_out = where.
void* alloc_near_gpu(size_t bytes, int gpu, int* node_out) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t len = (bytes + page - 1) / page * page;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    const int node = gpu_numa_node(gpu);
    if (node_out) *node_out = node;
    bool placed = false;
    cpu_set_t old_set, node_set;
    bool moved = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        placed = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, mask, 1024ul, 0u) == 0;
        if (!placed && node_cpus(node, &node_set) && sched_getaffinity(0, sizeof old_set, &old_set) == 0)
            moved = sched_setaffinity(0, sizeof node_set, &node_set) == 0;
    }
    for (size_t off = 0; off < len; off += page) static_cast<volatile char*>(p)[off] = 0;  // fault the pages in, here
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    return p;
}

void free_near_gpu(void* p, size_t bytes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    munmap(p, (bytes + page - 1) / page * page);
}

}  // namespace nsb
