// ddstore_b200/csrc/internal.h -- shared by the host translation units of libddstore_b200.so
#ifndef DDS_INTERNAL_H
#define DDS_INTERNAL_H
#include <string>

namespace dds_internal {
// record the calling thread's failure text and hand the code back (so call sites can `return fail(...)`)
int fail(int code, const std::string &detail = std::string());
void clear_error();
} // namespace dds_internal
#endif
