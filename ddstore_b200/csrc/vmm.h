// ddstore_b200/csrc/vmm.h -- CUDA VMM shard blocks + descriptor passing (see vmm.cpp)
#ifndef DDS_VMM_H
#define DDS_VMM_H
#include <stddef.h>

#include <string>
#include <vector>

#include "ddstore_b200.h"

namespace dds_vmm {

struct Block {
    void *ptr;
    size_t size;               // mapped size (multiple of the allocation granularity)
    unsigned long long handle; // CUmemGenericAllocationHandle
    int device;
    int fd;                    // exported POSIX fd (owner side), -1 otherwise
    bool mapped;
};

bool available(int device);
int alloc(int device, size_t bytes, Block *out);
int export_fd(Block *b);
int grant(const Block *b, int device); // let another device of THIS process read/write the block
int import_fd(int device, int fd, size_t size, Block *out);
void release(Block *b);
// my_fd < 0: nothing to export (sent explicitly); pids[r] = process id rank r published (sender verification)
int exchange_fds(dds_comm_t *comm, const std::string &tag, int my_fd, const std::vector<char> &want,
                 const std::vector<int> &pids, std::vector<int> *got);

} // namespace dds_vmm
#endif
