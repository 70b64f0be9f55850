// hixf_fuzz.cpp -- mutation fuzzer for the `.hixf` reader (taxor_b200/csrc/hixf_file.cpp), to be built with
// -fsanitize=address,undefined: a small valid index file is written, bytes are flipped / lengths are overwritten with
// huge values / the file is truncated or extended, and read_hixf must either parse it or return an error string --
// no crash, no sanitizer report, no allocation driven by an unchecked length.  An unmutated file must round-trip.
// usage: hixf_fuzz <tmp file> <iterations> <seed>
#include "../../taxor_b200/csrc/hixf_file.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

using namespace txr;

static TaxorIndexFile make_index(std::mt19937_64 &rng)
{
    TaxorIndexFile f;
    f.window_size = 20;
    f.kmer_size = 22;
    f.syncmer_size = 12;
    f.t_syncmer = 5;
    f.shape_size = 22;
    f.shape_bits = (1ull << 22) - 1;
    f.scaling = 1;
    const size_t n_ixf = 1 + rng() % 4, n_ub = 1 + rng() % 6;
    for (size_t u = 0; u < n_ub; ++u)
    {
        f.user_bin_filenames.push_back("/g/" + std::to_string(u) + ".fna");
        f.bin_path.push_back({"/g/" + std::to_string(u) + ".fna"});
        SpeciesRecord s;
        s.organism_name = "org" + std::to_string(u);
        s.accession_id = "GCF_" + std::to_string(u);
        s.taxid = std::to_string(100 + u);
        s.taxnames_string = "k__B;s__" + s.organism_name;
        s.taxid_string = "2;" + s.taxid;
        s.user_bin = u;
        s.seq_len = 1000 + u;
        f.species.push_back(s);
    }
    for (size_t i = 0; i < n_ixf; ++i)
    {
        IxfRecord x;
        x.seed = rng();
        x.bins = 1 + rng() % 5;
        x.tbins = 64;
        x.seg_len = 1 + rng() % 7;
        x.max_elems = 3;
        x.owned.resize(3 * x.seg_len * x.tbins);
        for (auto &b : x.owned)
            b = (uint8_t)rng();
        x.fp = x.owned.data();
        x.fp_len = x.owned.size();
        std::vector<int64_t> next(x.bins, (int64_t)i), ub(x.bins);
        for (auto &u : ub)
            u = (int64_t)(rng() % n_ub);
        f.next_ixf_id.push_back(next);
        f.ixf_bin_to_filename_position.push_back(ub);
        f.ixf.push_back(std::move(x));
    }
    for (auto &x : f.ixf)
        x.fp = x.owned.data();
    return f;
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
        TaxorIndexFile f = make_index(rng);
        const IxfRecordSpec &spec = IxfRecordSpec::candidates()[rng() % IxfRecordSpec::candidates().size()];
        const std::string werr = write_hixf(path, f, spec);
        if (!werr.empty())
        {
            fprintf(stderr, "write failed: %s\n", werr.c_str());
            return 1;
        }
        FILE *fp = fopen(path, "rb");
        std::vector<unsigned char> bytes;
        unsigned char buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0)
            bytes.insert(bytes.end(), buf, buf + n);
        fclose(fp);
        const int kind = (int)(rng() % 7);
        if (kind == 0)
        {
            // untouched: must parse with the spec it was written with and give the same geometry back
            TaxorIndexFile g;
            const std::string e = read_hixf(path, g, &spec, nullptr);
            if (!e.empty() || g.ixf.size() != f.ixf.size() || g.species.size() != f.species.size() ||
                g.ixf[0].seed != f.ixf[0].seed || g.ixf[0].fp_len != f.ixf[0].fp_len || memcmp(g.ixf[0].fp, f.ixf[0].fp, f.ixf[0].fp_len))
            {
                fprintf(stderr, "round trip failed at iteration %ld: %s\n", it, e.c_str());
                return 1;
            }
            ++accepted;
            continue;
        }
        if (kind == 1)
            for (int i = 0; i < 4; ++i)
                bytes[rng() % bytes.size()] = (unsigned char)rng();
        else if (kind == 2)
            bytes.resize(rng() % bytes.size());
        else if (kind == 3)
        {
            // overwrite an aligned u64 with a huge or odd value: lengths and counts are what an attacker would aim at
            const size_t at = (rng() % (bytes.size() / 8)) * 8 + (rng() % 2 ? 0 : 4);
            const uint64_t v = rng() % 3 == 0 ? ~0ULL : rng() % 3 == 1 ? (1ULL << (20 + rng() % 43)) : rng() % 5000;
            if (at + 8 <= bytes.size())
                memcpy(&bytes[at], &v, 8);
        }
        else if (kind == 4)
            bytes.insert(bytes.end(), rng() % 64, (unsigned char)rng());
        else if (kind == 5)
        {
            const size_t a = rng() % bytes.size(), b = std::min<size_t>(bytes.size(), a + rng() % 32);
            bytes.erase(bytes.begin() + (long)a, bytes.begin() + (long)b);
        }
        else
            bytes[rng() % std::min<size_t>(bytes.size(), 48)] ^= (unsigned char)(1u << (rng() % 8)); // header scalars
        fp = fopen(path, "wb");
        fwrite(bytes.data(), 1, bytes.size(), fp);
        fclose(fp);
        TaxorIndexFile g;
        const std::string e = read_hixf(path, g, rng() % 2 ? &spec : nullptr, nullptr);
        if (e.empty())
        {
            // whatever was accepted must be internally consistent enough to walk
            uint64_t sum = 0;
            for (auto &x : g.ixf)
            {
                if (x.fp_len != 3 * x.seg_len * x.tbins)
                    return 1;
                for (uint64_t i = 0; i < x.fp_len; i += 97)
                    sum += x.fp[i];
            }
            if (sum == 0xdeadbeefULL)
                printf(".");
            ++accepted;
        }
        else
            ++rejected;
    }
    printf("fuzz ok: %ld inputs, %ld accepted, %ld rejected\n", iters, accepted, rejected);
    return 0;
}
