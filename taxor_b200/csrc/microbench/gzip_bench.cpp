// gzip_bench.cpp -- host-side microbenchmark of the three ways `taxor search` can read a .gz file (no GPU involved):
// zlib's gzread (what the reference does through SeqAn3), the one-thread decoder of inflate_fast.cpp, and the multi-threaded
// reader of gzip_parallel.cpp.  Prints one JSON line; every path's output is checked against zlib's (length + CRC-32).
// usage: gzip_bench <file.gz> [threads ...]
#include "../gzip_parallel.hpp"
#include "../inflate_fast.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <vector>
#include <zlib.h>

using namespace txr;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 2)
        return 2;
    const int fd = open(argv[1], O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0)
        return 2;
    const uint8_t *d = static_cast<const uint8_t *>(mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0));
    std::vector<uint8_t> out(size_t(4) << 20);
    const int reps = 3;
    size_t want_n = 0;
    uint32_t want_crc = 0;
    double t_zlib = 1e30;
    for (int r = 0; r < reps; ++r)
    {
        const double t0 = now();
        gzFile g = gzopen(argv[1], "rb");
        gzbuffer(g, 1 << 20);
        size_t n_all = 0;
        uint32_t c = 0;
        for (;;)
        {
            const int n = gzread(g, out.data(), (unsigned)out.size());
            if (n <= 0)
                break;
            n_all += (size_t)n;
            if (!r)
                c = crc32_fast(c, out.data(), (size_t)n);
        }
        gzclose(g);
        t_zlib = std::min(t_zlib, now() - t0);
        if (!r)
            want_n = n_all, want_crc = c;
    }
    auto timed = [&](auto &&make_reader, bool &same) {
        double best = 1e30;
        for (int r = 0; r < reps; ++r)
        {
            const double t0 = now();
            auto reader = make_reader();
            size_t n_all = 0;
            uint32_t c = 0;
            for (;;)
            {
                const size_t n = reader->read(out.data(), out.size());
                if (!n)
                    break;
                n_all += n;
                if (!r)
                    c = crc32_fast(c, out.data(), n);
            }
            best = std::min(best, now() - t0);
            if (!r)
                same = n_all == want_n && c == want_crc;
        }
        return best;
    };
    bool same = false;
    const double t_serial = timed([&] { return std::make_unique<GzipStream>(d, (size_t)st.st_size); }, same);
    std::string par;
    bool all_same = same;
    for (int a = 2; a < argc; ++a)
    {
        const unsigned T = (unsigned)atoi(argv[a]);
        const double t = timed([&] { return std::make_unique<ParallelGzip>(d, (size_t)st.st_size, T); }, same);
        all_same = all_same && same;
        char buf[128];
        snprintf(buf, sizeof buf, "%s\"%u\": %.3f", par.empty() ? "" : ", ", T, want_n / 1e9 / t);
        par += buf;
    }
    printf("{\"file\": \"%s\", \"compressed_MB\": %.1f, \"inflated_MB\": %.1f, \"GBps_out\": {\"zlib_gzread\": %.3f, \"one_thread\": %.3f, "
           "\"threads\": {%s}}, \"best_of\": %d, \"outputs_equal_zlib\": %s}\n",
           argv[1], st.st_size / 1e6, want_n / 1e6, want_n / 1e9 / t_zlib, want_n / 1e9 / t_serial, par.c_str(), reps, all_same ? "true" : "false");
    return all_same ? 0 : 1;
}
