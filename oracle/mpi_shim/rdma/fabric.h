/* oracle/mpi_shim/rdma/fabric.h -- TEST INFRASTRUCTURE.
 * Just enough libfabric vocabulary for /root/reference/include/common.h:18-45 to parse.
 * The libfabric transport (method 1) is out of scope; its
 * three entry points are stubbed to abort() in oracle/ref_driver.cpp. */
#ifndef DDS_ORACLE_FABRIC_STUB_H
#define DDS_ORACLE_FABRIC_STUB_H
#include <stdint.h>
struct fi_context;
struct fi_info { uint64_t mode; };
struct fid_fabric;
struct fid_domain;
struct fid_ep;
struct fid_cq;
struct fid_av;
struct fid_mr;
typedef uint64_t fi_addr_t;
#define FI_LOCAL_MR (1ULL << 55)
#endif
