// tests/cpp/demo_ddstore.cpp -- the reference's C++ smoke (test/demo.cxx:20-37) re-expressed against
// include/ddstore_b200.hpp: every rank holds {1,2,3,4}+10*rank as nrows=2, disp=2 doubles and reads one row
// of its neighbour. Also exercises the exception types/texts of the C++ surface.
// usage: demo_ddstore <rank> <size> <shm-key> [device]
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

#include "ddstore_b200.hpp"

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    int rank = atoi(argv[1]), size = atoi(argv[2]);
    int device = argc > 4 ? atoi(argv[4]) : 0;
    dds_comm_t *comm = size > 1 ? dds_comm_shm(argv[3], rank, size) : dds_comm_self();
    if (!comm) {
        printf("comm failed: %s\n", dds_last_error());
        return 1;
    }
    try {
        DDStore ds(0, comm, device);
        const int N = 4;
        double buffer[N];
        for (int i = 0; i < N; i++) buffer[i] = i + 1 + 10 * rank;
        ds.add("var", buffer, 2, 2);
        double getbuf[4] = {0.0, 0.0, 0.0, 0.0};
        int start = (2 * (rank + 1)) % (2 * size) + 1;
        ds.get("var", start, 1, getbuf);
        printf("%d: start %d got %g %g %g %g\n", rank, start, getbuf[0], getbuf[1], getbuf[2], getbuf[3]);

        VarInfo_t vi;
        ds.query("var", vi);
        printf("%d: itemsize %d disp %d lenlist_last %ld\n", rank, vi.itemsize, vi.disp, vi.lenlist.back());
        try {
            float f[2];
            ds.get("var", 0, 1, f);
        } catch (const std::invalid_argument &e) { printf("%d: invalid_argument: %s\n", rank, e.what()); }
        try {
            ds.get("var", -1, 1, getbuf);
        } catch (const std::invalid_argument &e) { printf("%d: invalid_argument: %s\n", rank, e.what()); }
        try {
            ds.get("var", 2 * size, 1, getbuf);
        } catch (const std::invalid_argument &e) { printf("%d: invalid_argument: %s\n", rank, e.what()); }
        ds.epoch_begin();
        try {
            ds.epoch_begin();
        } catch (const std::logic_error &e) { printf("%d: logic_error: %s\n", rank, e.what()); }
        ds.epoch_end();
        try {
            ds.epoch_end();
        } catch (const std::logic_error &e) { printf("%d: logic_error: %s\n", rank, e.what()); }
        // batched form: all rows of the world, reversed
        std::vector<long> starts;
        for (long g = 2 * size - 1; g >= 0; g--) starts.push_back(g);
        std::vector<double> all(starts.size() * 2);
        long bytes = ds.get_batch<double>("var", starts.data(), nullptr, 1, (long)starts.size(), all.data(),
                                          (long)(all.size() * sizeof(double)));
        printf("%d: batch %ld bytes first %g last %g\n", rank, bytes, all.front(), all.back());
        ds.free();
    } catch (const std::exception &e) {
        printf("%d: FAILED: %s\n", rank, e.what());
        return 1;
    }
    dds_comm_free(comm);
    return 0;
}
