// profile_fuzz.cpp -- in-process mutation fuzzer for the head of `taxor profile` (taxor_b200/csrc/profile_ingest.cpp), meant
// to be built with -fsanitize=address,undefined: valid search-result files (the 10- and 6-column lines `taxor search`
// writes) are mutated (byte flips, truncation, dropped / doubled tabs, huge numbers, CRLF, NUL bytes), written to a file and
// pushed through txr_profile_add_file -> the three filter rounds -> both read-out calls.  A damaged file may be rejected with
// an error code; nothing may crash, throw through the C ABI or trip a sanitizer, and what is accepted must read out
// consistently (hit ranges ascending, as many 'R' lines as reads).
// usage: profile_fuzz <tmp file> <iterations> <seed>
#include "../../include/taxor_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

static std::string make_valid(std::mt19937_64 &rng)
{
    std::string s = "#QUERY_NAME\tACCESSION\tREFERENCE_NAME\tTAXID\tREF_LEN\tQUERY_LEN\tQHASH_COUNT\tQHASH_MATCH\tTAX_STR\tTAX_ID_STR\n";
    const int n_reads = 1 + (int)(rng() % 40), n_refs = 1 + (int)(rng() % 12);
    for (int r = 0; r < n_reads; ++r)
    {
        const std::string id = "read" + std::to_string(rng() % 4 == 0 ? rng() % n_reads : r) + (rng() % 5 == 0 ? " extra words" : "");
        const unsigned qlen = 100 + rng() % 20000, hashes = 1 + rng() % 900;
        const int n_hits = (int)(rng() % 5);
        if (!n_hits)
        {
            s += id + "\t-\t-\t-\t-\t" + std::to_string(qlen) + "\n";
            continue;
        }
        for (int h = 0; h < n_hits; ++h)
        {
            const int ref = (int)(rng() % n_refs);
            s += id + "\tGCF_" + std::to_string(ref) + "\torganism " + std::to_string(ref) + "\t" + std::to_string(1000 + ref) + "\t" +
                 std::to_string(1000000 + 1000 * ref) + "\t" + std::to_string(qlen) + "\t" + std::to_string(hashes) + "\t" +
                 std::to_string(rng() % (hashes + 1)) + "\tBacteria;sp" + std::to_string(ref) + "\t2;" + std::to_string(1000 + ref) + "\n";
        }
    }
    return s;
}

static void mutate(std::string &s, std::mt19937_64 &rng)
{
    if (s.empty())
        return;
    switch (rng() % 9)
    {
    case 0:
        for (int i = 0; i < 4; ++i)
            s[rng() % s.size()] = (char)(rng() % 256);
        break;
    case 1:
        s.resize(rng() % s.size());
        break;
    case 2: // drop some tabs
        for (int i = 0; i < 3; ++i)
        {
            const size_t p = s.find('\t', rng() % s.size());
            if (p != std::string::npos)
                s.erase(p, 1);
        }
        break;
    case 3: // double some tabs (empty columns)
        for (int i = 0; i < 3; ++i)
        {
            const size_t p = s.find('\t', rng() % s.size());
            if (p != std::string::npos)
                s.insert(p, "\t");
        }
        break;
    case 4: // a number far beyond 64 bits, a negative one, a non-number
    {
        const char *junk[] = {"99999999999999999999999999999", "-5", "1e9", "", " 12", "0x10"};
        const size_t p = s.find('\t', rng() % s.size());
        if (p != std::string::npos)
            s.insert(p + 1, junk[rng() % 6]);
        break;
    }
    case 5: // CRLF
        for (size_t p = 0; (p = s.find('\n', p)) != std::string::npos; p += 2)
            s.insert(p, "\r");
        break;
    case 6:
        s.insert(rng() % s.size(), std::string(1 + rng() % 3, '\0'));
        break;
    case 7:
        s.insert(rng() % s.size(), std::string(rng() % 30, "\t\n-;"[rng() % 4]));
        break;
    default: // keep valid
        break;
    }
}

int main(int argc, char **argv)
{
    if (argc < 4)
        return 2;
    const char *path = argv[1];
    const long iters = atol(argv[2]);
    std::mt19937_64 rng((uint64_t)atoll(argv[3]));
    long accepted = 0, rejected = 0;
    for (long it = 0; it < iters; ++it)
    {
        std::string text = make_valid(rng);
        const int rounds = (int)(rng() % 3);
        for (int m = 0; m < rounds; ++m)
            mutate(text, rng);
        FILE *f = fopen(path, "wb");
        if (!f)
            return 3;
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);

        txr_profile *p = nullptr;
        if (txr_profile_create(&p) != TXR_OK)
            return 4;
        const int rc = txr_profile_add_file(p, path);
        if (rc != TXR_OK)
        {
            ++rejected;
            if (!*txr_profile_last_error())
            {
                printf("rejected without a message (iteration %ld)\n", it);
                return 1;
            }
            txr_profile_destroy(p);
            continue;
        }
        ++accepted;
        if (rng() & 1) // the same file twice: the second pass appends to the reads of the first
            (void)txr_profile_add_file(p, path);
        for (int r = 0; r <= 3; ++r)
        {
            if (txr_profile_filter(p, r) != TXR_OK)
            {
                printf("filter round %d failed: %s\n", r, txr_profile_last_error());
                return 1;
            }
            txr_profile_view v;
            const char *dump = nullptr;
            uint64_t len = 0;
            if (txr_profile_get(p, &v) != TXR_OK || txr_profile_text(p, &dump, &len) != TXR_OK)
                return 1;
            uint64_t lines_r = 0;
            for (uint64_t i = 0; i + 1 < len; ++i)
                lines_r += (i == 0 || dump[i - 1] == '\n') && dump[i] == 'R' && dump[i + 1] == '\t';
            for (uint64_t i = 0; i < v.n_reads; ++i)
                if (v.hit_begin[i] > v.hit_begin[i + 1] || !v.read_id[i])
                {
                    printf("inconsistent view (iteration %ld)\n", it);
                    return 1;
                }
            if (lines_r < v.n_reads) // a read id with an embedded line break can add 'R'-looking lines, never remove one
            {
                printf("dump has %llu reads, view %llu (iteration %ld)\n", (unsigned long long)lines_r, (unsigned long long)v.n_reads, it);
                return 1;
            }
        }
        // adding after the filter is a state error, not a crash
        if (txr_profile_add_file(p, path) == TXR_OK)
        {
            printf("add after filter accepted\n");
            return 1;
        }
        txr_profile_destroy(p);
    }
    printf("fuzz ok: %ld accepted, %ld rejected\n", accepted, rejected);
    return 0;
}
