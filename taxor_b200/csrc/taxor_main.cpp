// taxor_main.cpp -- drop-in `taxor search` command (src/main/main.cpp:51-88 + src/main/taxor_search.cpp) on top of
// the CUDA library.  Same sub-command, options, validators, stdout banners, result-file format and exit codes as
// the reference; `build` and `profile` stay with the reference binary (CPU, out of scope).
// Extra, optional knobs (ignored by the reference): --gpus <n|list>, --ixf-record <field order>.
#include "../../include/taxor_b200.h"
#include "hixf_file.hpp"
#include "ingest.hpp"
#include "threshold.hpp"

#include <algorithm>
#include <atomic>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <sys/resource.h>
#include <sys/stat.h>
#include <thread>
#include <vector>

using namespace txr;

namespace
{
struct Config // src/main/taxor_search_configuration.hpp:8-19
{
    std::string index_file, query_file, report_file;
    double threshold{-1.0};
    double error_rate{0.04};
    unsigned threads{1};
    bool threads_given{false};
    std::string gpus;       // extension
    std::string ixf_record; // extension
    std::string ixf_scheme; // extension: probe arithmetic of the index (hixf_file.hpp IxfSchemeSpec), default the prototype's
    bool debug{false};
};

struct ParserError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

double cputime() // main.cpp:37-42
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}
long peak_rss() // main.cpp:44-49
{
    struct rusage r;
    getrusage(RUSAGE_SELF, &r);
    return r.ru_maxrss * 1024;
}

std::vector<std::string> str_split(const std::string &s, char delim) // taxor_search.cpp:82-95
{
    std::vector<std::string> out;
    std::stringstream ss(s);
    std::string seg;
    while (std::getline(ss, seg, delim))
        out.push_back(seg);
    return out;
}

bool file_exists(const std::string &p)
{
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

void print_help()
{
    std::cout << "taxor-search - Queries files of DNA sequences against a list of HIXF index files\n"
                 "=================================================================================\n\n"
                 "DESCRIPTION\n    Query sequences against the taxor HIXF index structure\n\n"
                 "OPTIONS\n\n  Main options:\n"
                 "    --index-file (std::string)\n          taxor index file containing HIXF index and reference sequence information\n"
                 "    --query-file (std::string)\n          file containing sequences to query against the index Default: .\n"
                 "    --output-file (std::string)\n          A file name for the resulting output. Default: .\n"
                 "    --threads (unsigned 8 bit integer)\n          The number of threads to use. Default: 1. Value must be in range [1,32].\n"
                 "    --percentage (double)\n          If set, this threshold is used instead of the k-mer/syncmer models. Default: -1. Value must be in range\n          [0,1].\n"
                 "    --error-rate (double)\n          Expected error rate of reads that will be queried Default: 0.04. Value must be in range [0,1].\n\n"
                 "  B200 options (extensions of this implementation):\n"
                 "    --gpus (std::string)\n          Number of GPUs or comma-separated device list. Default: all visible devices.\n"
                 "    --ixf-record (std::string)\n          Field order of the interleaved XOR filter records in the index file (see INTEGRATION.md).\n"
                 "    --ixf-scheme (std::string)\n          Probe arithmetic of the interleaved XOR filters: xor3 | fuse3 [:mix=add|xor,fp=fold32|low8|high8,\n          rot=R1/R2,layout=slot|bin]. Default: xor3 (the prototype's arithmetic, see INTEGRATION.md).\n\n"
                 "VERSION\n    taxor-search version: 0.2.0 (taxor_b200)\n";
}

double parse_double(const std::string &opt, const std::string &v, double lo, double hi)
{
    char *end = nullptr;
    const double d = strtod(v.c_str(), &end);
    if (v.empty() || *end)
        throw ParserError("Value parse failed for --" + opt + ": Argument " + v + " could not be parsed as type double.");
    if (d < lo || d > hi)
    {
        std::ostringstream os;
        os << "Validation failed for option --" << opt << ": Value " << d << " is not in range [" << lo << "," << hi << "].";
        throw ParserError(os.str());
    }
    return d;
}

void parse_args(int argc, char const **argv, Config &c) // taxor_search.cpp:32-80
{
    bool have_index = false;
    for (int i = 0; i < argc; ++i)
    {
        std::string a = argv[i], v;
        if (a == "-h" || a == "--help" || a == "-hh" || a == "--advanced-help")
        {
            print_help();
            exit(0);
        }
        if (a == "--version")
        {
            std::cout << "taxor-search version: 0.2.0 (taxor_b200)\n";
            exit(0);
        }
        if (a == "--output-verbose-statistics" || a == "--debug") // hidden flags, no effect on search (:69-79)
        {
            c.debug = c.debug || a == "--debug"; // here: also prints which record order / filter scheme the index was read with
            continue;
        }
        if (a.rfind("--", 0) != 0)
            throw ParserError("Too many arguments provided. Please see -h/--help for more information.");
        const size_t eq = a.find('=');
        std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
        if (eq != std::string::npos)
            v = a.substr(eq + 1);
        else
        {
            if (i + 1 >= argc)
                throw ParserError("Missing value for option --" + name);
            v = argv[++i];
        }
        if (name == "index-file")
        {
            c.index_file = v;
            have_index = true;
        }
        else if (name == "query-file")
            c.query_file = v;
        else if (name == "output-file")
            c.report_file = v;
        else if (name == "threads")
        {
            char *end = nullptr;
            const long t = strtol(v.c_str(), &end, 10);
            if (v.empty() || *end)
                throw ParserError("Value parse failed for --threads: Argument " + v + " could not be parsed as type unsigned 8 bit integer.");
            if (t < 1 || t > 32)
                throw ParserError("Validation failed for option --threads: Value " + std::to_string(t) + " is not in range [1,32].");
            c.threads = (unsigned)t;
            c.threads_given = true;
        }
        else if (name == "percentage")
            c.threshold = parse_double(name, v, 0.0, 1.0);
        else if (name == "error-rate")
            c.error_rate = parse_double(name, v, 0.0, 1.0);
        else if (name == "gpus")
            c.gpus = v;
        else if (name == "ixf-record")
            c.ixf_record = v;
        else if (name == "ixf-scheme")
        {
            IxfSchemeSpec probe;
            std::string err;
            if (!IxfSchemeSpec::parse(v, probe, err))
                throw ParserError("Validation failed for option --ixf-scheme: " + err);
            c.ixf_scheme = v;
        }
        else
            throw ParserError("Unknown option --" + name + ". In case this is meant to be a non-option/argument/parameter, please specify the start of non-options with '--'. See -h/--help for program information.");
    }
    if (!have_index)
        throw ParserError("Option --index-file is required but not set.");
}

IxfSchemeSpec scheme_of(const Config &cfg)
{
    IxfSchemeSpec sc;
    std::string err;
    if (!cfg.ixf_scheme.empty())
        IxfSchemeSpec::parse(cfg.ixf_scheme, sc, err); // validated by parse_args
    return sc;
}

std::string load_index_file(const std::string &path, const Config &cfg, TaxorIndexFile &idx, HixfReadReport *report = nullptr)
{
    const IxfSchemeSpec sc = scheme_of(cfg);
    if (cfg.ixf_record.empty())
        return read_hixf(path, idx, nullptr, nullptr, &sc, report);
    const IxfRecordSpec spec = IxfRecordSpec::parse(cfg.ixf_record);
    return read_hixf(path, idx, &spec, nullptr, &sc, report);
}

// ---- one GPU worker: owns a context with the index in HBM ----
struct Chunk
{
    size_t seq{0};
    size_t n{0};                    // reads in this chunk
    std::vector<std::string> ids;   // pre-sized to kChunkReads: filled by the pack threads, never reallocated
    std::vector<uint32_t> len;
    std::vector<uint64_t> word_off;
    uint64_t *words{nullptr};
    size_t words_cap{0}, words_used{0};
    std::atomic<int> pending{0};    // pack tasks still running + 1 while the reader may add more
    std::string text; // formatted result lines
};

struct RawBuf // streaming path (gzip): a stretch of the input that starts at a record boundary
{
    std::vector<char> data;
    std::vector<RecordRef> recs;
};

struct Segment : SegmentScan // mapped path: the records that start inside one byte range of the file
{
    bool done{false};
};

template <typename T> class Channel
{
public:
    void push(T v)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back(std::move(v));
        }
        cv_.notify_one();
    }
    bool pop(T &v)
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty())
            return false;
        v = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    void close()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            closed_ = true;
        }
        cv_.notify_all();
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool closed_{false};
};

std::vector<int> pick_devices(const std::string &spec)
{
    std::vector<int> dev;
    if (spec.find(',') != std::string::npos)
    {
        for (auto &s : str_split(spec, ','))
            dev.push_back(atoi(s.c_str()));
        return dev;
    }
    int n = spec.empty() ? 64 : atoi(spec.c_str());
    for (int d = 0; d < n; ++d)
    {
        txr_ctx *probe = nullptr;
        if (txr_ctx_create(d, &probe) != TXR_OK)
            break;
        txr_ctx_destroy(probe);
        dev.push_back(d);
    }
    return dev;
}

// result lines of one chunk, taxor_search.cpp:266-306
void format_chunk(Chunk &ch, const txr_result &res, const TaxorIndexFile &idx, const std::map<size_t, size_t> &user_bin_index)
{
    std::string &out = ch.text;
    out.clear();
    for (size_t r = 0; r < ch.n; ++r)
    {
        const std::string &id = ch.ids[r];
        const uint64_t b = res.hit_begin[r], e = res.hit_begin[r + 1];
        if (b == e) // :268-273
        {
            out += id;
            out += '\t';
            out += "-\t-\t-\t-\t";
            out += std::to_string(ch.len[r]);
            out += "\n";
            continue;
        }
        for (uint64_t i = b; i < e; ++i)
        {
            if (!res.keep[i]) // :285-286 (count < 0.8 * max_count)
                continue;
            auto it = user_bin_index.find((size_t)res.user_bin[i]);
            const SpeciesRecord &sp = idx.species.at(it == user_bin_index.end() ? 0 : it->second);
            out += id;
            out += '\t';
            out += sp.accession_id;
            out += '\t';
            out += sp.organism_name;
            out += '\t';
            out += sp.taxid;
            out += '\t';
            out += std::to_string(sp.seq_len);
            out += '\t';
            out += std::to_string(ch.len[r]);
            out += '\t';
            out += std::to_string(res.hash_count[r]);
            out += '\t';
            out += std::to_string(res.count[i]);
            out += '\t';
            out += sp.taxnames_string;
            out += '\t';
            out += sp.taxid_string;
            out += '\n';
        }
    }
}

constexpr size_t kChunkReads = 65536;
constexpr size_t kChunkBases = 400u << 20;
constexpr size_t kRawTarget = 32u << 20; // bytes of input per pack task

// search_single, taxor_search.cpp:153-338
void search_single(const Config &cfg, const std::string &query, const std::string &index_path, std::ofstream &out,
                   const std::vector<int> &devices)
{
    const bool timing = getenv("TAXOR_TIMING") != nullptr; // stderr: wall time of the load / upload / search phases
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count(); };
    TaxorIndexFile idx;
    std::string load_error;
    HixfReadReport read_report;
    std::thread loader([&] { load_error = load_index_file(index_path, cfg, idx, &read_report); }); // the async cereal_worker (:162-180)

    MappedFile mapped(query);
    if (!mapped.ok())
    {
        loader.join();
        throw std::runtime_error("cannot open query file " + query);
    }
    loader.join();
    if (!load_error.empty())
        throw std::runtime_error(load_error);
    // the record order and the probe arithmetic of the IXFs are defined by a SeqAn3 fork that is not available:
    // say which ones were used when asked, and say so unasked when the file left the choice open
    if (!read_report.note.empty())
        std::cerr << "[TAXOR SEARCH WARNING] " << index_path << ": " << read_report.note << std::endl;
    if (cfg.debug)
        std::cerr << "[taxor debug] " << index_path << ": IXF record order [" << read_report.used.str() << "], filter scheme "
                  << scheme_of(cfg).str() << ", " << read_report.tiling_candidates << " order(s) tile the file, "
                  << read_report.capacity_consistent << " consistent with max_elems" << std::endl;

    Thresholder thresholder(idx.window_size, idx.kmer_size, cfg.threshold, cfg.error_rate, idx.use_syncmer);
    std::cout << thresholder.banner();
    if (thresholder.kind() == ThresholdKind::percentage)
        std::cout << "\t" << cfg.threshold; // threshold.hpp:32
    std::cout << std::endl << std::flush;
    std::map<size_t, size_t> user_bin_index; // :172-178
    for (size_t i = 0; i < idx.species.size(); ++i)
        user_bin_index.emplace(idx.species[i].user_bin, i);

    // index view shared by all contexts
    std::vector<txr_ixf_view> views(idx.ixf.size());
    std::vector<uint64_t> bin_off(idx.ixf.size() + 1, 0);
    std::vector<int64_t> next_flat, ub_flat;
    for (size_t i = 0; i < idx.ixf.size(); ++i)
    {
        views[i] = txr_ixf_view{idx.ixf[i].seed, idx.ixf[i].bins, idx.ixf[i].tbins, idx.ixf[i].seg_len, idx.ixf[i].fp, idx.ixf[i].rows};
        next_flat.insert(next_flat.end(), idx.next_ixf_id[i].begin(), idx.next_ixf_id[i].end());
        ub_flat.insert(ub_flat.end(), idx.ixf_bin_to_filename_position[i].begin(), idx.ixf_bin_to_filename_position[i].end());
        bin_off[i + 1] = next_flat.size();
    }
    const IxfSchemeSpec sch = scheme_of(cfg);
    const txr_ixf_scheme scheme{sch.slots, sch.mix, sch.fingerprint, sch.rot1, sch.rot2, sch.layout};
    txr_hixf_view hv{idx.ixf.size(), views.data(), bin_off.data(), next_flat.data(), ub_flat.data(), idx.user_bin_filenames.size(), &scheme};
    txr_params par{};
    par.kmer_size = idx.kmer_size;
    par.syncmer_size = idx.syncmer_size;
    par.t_syncmer = idx.t_syncmer;
    par.use_syncmer = idx.use_syncmer;
    par.window_size = (uint32_t)idx.window_size;
    par.scaling = idx.scaling;
    par.percentage = cfg.threshold;
    par.error_rate = cfg.error_rate;

    const double t_load = since(t_start);
    const auto t_up = std::chrono::steady_clock::now();
    // one upload over PCIe (staged, multi-threaded), then the other GPUs are filled device to device in a doubling
    // tree (round r: every context that has the index clones it into one that has not -- NVLink between peers), instead
    // of pushing the same 10-100 GB through the host once per GPU
    // (creating a CUDA context costs a few hundred ms per device: the contexts of the other GPUs come up on their own
    // threads while the first one takes the upload)
    std::vector<txr_ctx *> ctxs(devices.size(), nullptr);
    std::vector<std::string> ctx_err(devices.size());
    auto make_ctx = [&](size_t i) {
        if (txr_ctx_create(devices[i], &ctxs[i]) != TXR_OK)
            ctx_err[i] = txr_last_error();
    };
    make_ctx(0);
    // The pinned chunk pool (scan -> pack jobs -> GPU workers -> ordered writer) is allocated on its own thread from here on:
    // pinning ~100 MB per chunk costs tens of ms each, which used to sit between the upload and the first parsed byte; now the
    // buffers come up while the index is uploaded, and parsing starts as soon as ONE exists.
    const size_t n_chunks = 2 * devices.size() + 2;
    std::vector<Chunk> pool(n_chunks);
    Channel<Chunk *> free_q, work_q, done_q;
    std::string pool_error;
    double t_pool = 0;
    std::thread pool_thread;
    struct JoinOnExit // the allocation thread never outlives `pool`, whatever is thrown below
    {
        std::thread &t;
        ~JoinOnExit()
        {
            if (t.joinable())
                t.join();
        }
    } pool_join{pool_thread};
    if (ctx_err[0].empty())
        pool_thread = std::thread([&] {
            const auto t0 = std::chrono::steady_clock::now();
            for (auto &ch : pool)
            {
                ch.words_cap = kChunkBases / 32 + 2 * kChunkReads + 64;
                ch.words = static_cast<uint64_t *>(txr_ctx_host_alloc(ctxs[0], ch.words_cap * 8));
                if (!ch.words)
                {
                    pool_error = txr_last_error();
                    free_q.close(); // the assembler finds the pool dry instead of waiting for ever
                    return;
                }
                ch.ids.resize(kChunkReads);
                ch.len.resize(kChunkReads);
                ch.word_off.resize(kChunkReads);
                free_q.push(&ch);
            }
            t_pool = since(t0);
        });
    std::vector<std::thread> ctx_threads;
    for (size_t i = 1; i < devices.size(); ++i)
        ctx_threads.emplace_back(make_ctx, i);
    std::string up_err;
    if (ctx_err[0].empty() && txr_index_upload(ctxs[0], &hv) != TXR_OK)
        up_err = txr_last_error();
    for (auto &t : ctx_threads)
        t.join();
    for (size_t i = 0; i < devices.size(); ++i)
        if (!ctx_err[i].empty())
            throw std::runtime_error(std::string("GPU ") + std::to_string(devices[i]) + ": " + ctx_err[i]);
    if (!up_err.empty())
        throw std::runtime_error(std::string("GPU ") + std::to_string(devices[0]) + ": " + up_err);
    for (size_t have = 1; have < ctxs.size(); have *= 2)
    {
        std::vector<std::thread> th;
        std::vector<std::string> errs(ctxs.size());
        for (size_t i = 0; i < have && have + i < ctxs.size(); ++i)
            th.emplace_back([&, i] {
                if (txr_index_clone(ctxs[have + i], ctxs[i]) != TXR_OK)
                    errs[have + i] = txr_last_error();
            });
        for (auto &t : th)
            t.join();
        for (size_t i = 0; i < errs.size(); ++i)
            if (!errs[i].empty())
                throw std::runtime_error(std::string("GPU ") + std::to_string(devices[i]) + ": " + errs[i]);
    }
    for (size_t i = 0; i < ctxs.size(); ++i)
        if (txr_params_set(ctxs[i], &par) != TXR_OK)
            throw std::runtime_error(std::string("GPU ") + std::to_string(devices[i]) + ": " + txr_last_error());

    const double t_upload = since(t_up);
    const auto t_search = std::chrono::steady_clock::now();
    // chunk pool (pinned): scan -> pack jobs -> GPU workers -> ordered writer
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned n_pack = cfg.threads_given ? cfg.threads : std::min(16u, hw);
#ifdef _OPENMP
    omp_set_num_threads((int)n_pack); // block-parallel BGZF inflation inside the scanner obeys --threads as well
#endif
    Channel<std::function<void()>> jobs;
    // where the phase goes (TAXOR_TIMING=1): seconds the assembler waited for a scanned segment / for a free chunk, seconds the
    // GPU workers spent in txr_search / formatting, seconds the writer spent writing
    double w_seg = 0, w_chunk = 0, w_write = 0, w_stream = 0;
    std::atomic<uint64_t> us_search{0}, us_format{0}, us_scan{0}, us_pack{0};
    std::string worker_error;
    std::mutex err_m;
    std::atomic<uint64_t> n_reads_total{0}, n_reads_hit{0};
    std::atomic<bool> failed{false}; // a parse or search error: everything still queued drains without further work
    std::vector<std::thread> workers;
    for (txr_ctx *c : ctxs)
        workers.emplace_back([&, c] {
            // the buffers a chunk-sized search needs, allocated while the first chunk is still being parsed and packed
            txr_ctx_reserve(c, kChunkReads, kChunkBases);
            Chunk *ch;
            while (work_q.pop(ch))
            {
                txr_result res{};
                if (failed.load()) // after the first failure no context is searched again; the chunks only circulate
                    ch->text.clear();
                else if ([&] {
                             const auto t0 = std::chrono::steady_clock::now();
                             const int rc = txr_search(c, ch->words, ch->word_off.data(), ch->len.data(), ch->n, &res);
                             us_search += (uint64_t)(since(t0) * 1e6);
                             return rc;
                         }() != TXR_OK)
                {
                    std::lock_guard<std::mutex> l(err_m);
                    if (worker_error.empty())
                        worker_error = txr_last_error();
                    ch->text.clear();
                    failed = true;
                }
                else
                {
                    const auto t0 = std::chrono::steady_clock::now();
                    format_chunk(*ch, res, idx, user_bin_index);
                    us_format += (uint64_t)(since(t0) * 1e6);
                    uint64_t hit = 0;
                    for (size_t r = 0; r < ch->n; ++r)
                        hit += res.hit_begin[r + 1] > res.hit_begin[r];
                    n_reads_total += ch->n;
                    n_reads_hit += hit;
                }
                done_q.push(ch);
            }
        });
    std::thread writer([&] {
        std::map<size_t, Chunk *> pending;
        size_t next = 0;
        Chunk *ch;
        while (done_q.pop(ch))
        {
            pending[ch->seq] = ch;
            while (!pending.empty() && pending.begin()->first == next)
            {
                Chunk *w = pending.begin()->second;
                pending.erase(pending.begin());
                const auto t0 = std::chrono::steady_clock::now();
                out << w->text; // sync_out::write (:311): here in read order
                w_write += since(t0);
                ++next;
                free_q.push(w);
            }
        }
    });

    // job threads: segment scans (mapped files) and packing (IUPAC -> dna4 -> 2 bit straight into the pinned chunk, ids)
    std::string parse_error;
    auto fail = [&](const std::string &msg) {
        std::lock_guard<std::mutex> l(err_m);
        if (parse_error.empty())
            parse_error = msg;
        failed = true;
    };
    // state the jobs point into: declared before the threads so that it outlives them on every path
    std::vector<RawBuf> raw_pool(n_pack + 2);
    Channel<RawBuf *> raw_free;
    std::mutex seg_m;
    std::condition_variable seg_cv;
    std::vector<std::thread> job_threads;
    for (unsigned t = 0; t < n_pack; ++t)
        job_threads.emplace_back([&] {
            std::function<void()> job;
            while (jobs.pop(job))
                job();
        });
    auto pack_job = [&](std::shared_ptr<const void> hold, const char *base, const RecordRef *recs, uint32_t count, Chunk *chp, uint32_t first_read) {
        jobs.push([&, hold, base, recs, count, chp, first_read] {
            Chunk &ch = *chp;
            thread_local std::string scratch;
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t k = 0; k < count && !failed.load(std::memory_order_relaxed); ++k)
            {
                const RecordRef &r = recs[k];
                const uint32_t idx = first_read + k;
                ch.ids[idx].assign(base + r.id_off, r.id_len);
                const char *bases = base + r.seq_off;
                if (!r.single_line)
                {
                    join_record(base, r, scratch);
                    bases = scratch.data();
                }
                if (txr_pack_2bit(bases, r.seq_len, ch.words + ch.word_off[idx]) != TXR_OK)
                {
                    // blanks (and, in FASTA, digits) inside sequence lines are not bases: SeqAn3 filters them before the
                    // alphabet check.  The record keeps its (larger) word reservation; only its length shrinks.
                    clean_record(base, r, scratch);
                    if (scratch.size() == r.seq_len || txr_pack_2bit(scratch.data(), scratch.size(), ch.words + ch.word_off[idx]) != TXR_OK)
                        fail("read '" + ch.ids[idx] + "': " + txr_last_error());
                    else
                        ch.len[idx] = (uint32_t)scratch.size();
                }
            }
            us_pack += (uint64_t)(since(t0) * 1e6);
            if (ch.pending.fetch_sub(1) == 1)
                work_q.push(&ch);
        });
    };

    // assembler (this thread): records -> chunk slots in file order; chunks are taken and sealed in order
    size_t seq = 0;
    Chunk *cur = nullptr;
    size_t cur_bases = 0;
    auto seal = [&] {
        if (!cur)
            return;
        cur->seq = seq++;
        if (cur->pending.fetch_sub(1) == 1)
            work_q.push(cur);
        cur = nullptr;
    };
    auto assemble = [&](const std::shared_ptr<const void> &hold, const char *base, const std::vector<RecordRef> &recs) {
        size_t i = 0;
        const size_t n_rec = recs.size();
        while (i < n_rec)
        {
            if (!cur)
            {
                const auto t0 = std::chrono::steady_clock::now();
                if (!free_q.pop(cur)) // only a failed pool allocation closes this queue
                {
                    cur = nullptr;
                    throw std::runtime_error(pool_error.empty() ? "chunk pool exhausted" : pool_error);
                }
                w_chunk += since(t0);
                cur->n = 0;
                cur->words_used = 0;
                cur->pending = 1;
                cur_bases = 0;
            }
            size_t j = i;
            while (j < n_rec)
            {
                const uint64_t L = recs[j].seq_len, nw = txr_packed_words(L);
                if (cur->n >= kChunkReads || (cur->n && cur_bases + L > kChunkBases) || cur->words_used + nw > cur->words_cap)
                    break;
                cur->len[cur->n] = (uint32_t)L;
                cur->word_off[cur->n] = cur->words_used;
                cur->words_used += nw;
                cur_bases += L;
                ++cur->n;
                ++j;
            }
            if (j == i)
            {
                if (cur->n) // full: hand it over and start the next one
                {
                    seal();
                    continue;
                }
                // a single sequence larger than a chunk: grow this (empty, job-free) chunk's buffer
                const uint64_t nw = txr_packed_words(recs[i].seq_len);
                txr_host_free(cur->words);
                cur->words_cap = nw + 64;
                cur->words = static_cast<uint64_t *>(txr_host_alloc(cur->words_cap * 8));
                if (!cur->words)
                    throw std::runtime_error(txr_last_error());
                continue;
            }
            cur->pending.fetch_add(1);
            pack_job(hold, base, recs.data() + i, (uint32_t)(j - i), cur, (uint32_t)(cur->n - (j - i)));
            i = j;
        }
    };

    try
    {
        if (mapped.ok() && !mapped.gzip())
        {
            // mapped file: byte segments scanned by the job threads; every guess of a segment's first record is checked
            // against the exact end of the segment before it, and a wrong one is rescanned here
            const char *data = mapped.data();
            const size_t size = mapped.size();
            size_t first = 0;
            while (first < size && (data[first] == '\n' || data[first] == '\r'))
                ++first;
            if (first < size && data[first] != '>' && data[first] != '@')
                throw std::runtime_error("sequence file: record does not start with '>' or '@'");
            const char marker = first < size ? data[first] : '>';
            const size_t n_seg = (size + kRawTarget - 1) / kRawTarget;
            std::vector<std::shared_ptr<Segment>> segs(n_seg);
            size_t dispatched = 0;
            auto dispatch = [&](size_t upto) {
                for (; dispatched < std::min(upto, n_seg); ++dispatched)
                {
                    auto sg = std::make_shared<Segment>();
                    segs[dispatched] = sg;
                    const size_t lo = dispatched * kRawTarget, hi = std::min(size, lo + kRawTarget);
                    jobs.push([&seg_m, &seg_cv, &us_scan, sg, lo, hi, data, size, first, marker] {
                        const auto t0 = std::chrono::steady_clock::now();
                        scan_byte_range(data, size, first, marker, lo, hi, *sg);
                        us_scan += (uint64_t)(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6);
                        {
                            std::lock_guard<std::mutex> l(seg_m);
                            sg->done = true;
                        }
                        seg_cv.notify_all();
                    });
                }
            };
            size_t expected = first;
            for (size_t k = 0; k < n_seg && !failed; ++k)
            {
                dispatch(k + 2 * n_pack + 2);
                std::shared_ptr<Segment> sg = segs[k];
                {
                    const auto t0 = std::chrono::steady_clock::now();
                    std::unique_lock<std::mutex> l(seg_m);
                    seg_cv.wait(l, [&] { return sg->done; });
                    w_seg += since(t0);
                }
                segs[k].reset();
                const size_t hi = std::min(size, (k + 1) * kRawTarget);
                expected = accept_byte_range(data, size, expected, hi, *sg); // a wrong guess is rescanned exactly here
                assemble(sg, data + sg->begin, sg->recs);
            }
        }
        else
        {
            // gzip (or unmappable) input: one streaming scanner, raw buffers recycled through a pool
            RecordScanner fin(query);
            if (!fin.ok())
                throw std::runtime_error(fin.open_error().empty() ? "cannot open query file " + query : query + ": " + fin.open_error());
            for (auto &rb : raw_pool)
                raw_free.push(&rb);
            while (!failed)
            {
                RawBuf *rb = nullptr;
                raw_free.pop(rb);
                const auto t0 = std::chrono::steady_clock::now();
                const bool more = fin.next(rb->data, rb->recs, kRawTarget);
                w_stream += since(t0);
                if (!more)
                    break;
                std::shared_ptr<RawBuf> hold(rb, [&raw_free](RawBuf *p) { raw_free.push(p); });
                assemble(hold, rb->data.data(), rb->recs);
            }
        }
        seal();
    }
    catch (std::exception const &e)
    {
        fail(e.what());
        seal();
    }
    if (pool_thread.joinable())
        pool_thread.join();
    jobs.close();
    for (auto &t : job_threads)
        t.join();
    work_q.close();
    for (auto &t : workers)
        t.join();
    done_q.close();
    writer.join();
    if (timing)
    {
        uint64_t fp_bytes = 0;
        for (auto &x : idx.ixf)
            fp_bytes += 3 * x.seg_len * x.tbins;
        std::cerr << "[taxor timing] index load " << t_load << " s, upload to " << ctxs.size() << " GPU(s) " << t_upload << " s ("
                  << fp_bytes / 1e9 << " GB), ingest+search+write " << since(t_search) << " s, " << seq << " chunks, " << n_pack
                  << " pack threads\n"
                  << "[taxor timing]   chunk pool " << t_pool << " s (on its own thread, from the start of the upload); assembler waited " << w_seg << " s for scans, " << w_chunk
                  << " s for a free chunk, spent " << w_stream << " s in the streaming reader (compressed input); job threads: scan " << us_scan / 1e6 << " s, pack " << us_pack / 1e6 << " s (summed over threads); GPU workers: search "
                  << us_search / 1e6 << " s, format " << us_format / 1e6 << " s; writer " << w_write << " s\n";
    }
    // nothing at all matched: with a reference-built index that is what a wrong guess of the (unpinned) record order or
    // filter arithmetic looks like -- every probe misses silently.  Say it loudly instead.
    if (n_reads_total.load() >= 1000 && n_reads_hit.load() == 0)
        std::cerr << "[TAXOR SEARCH WARNING] none of the " << n_reads_total.load() << " reads matched any reference. If this index was "
                     "written by the reference `taxor build`, its interleaved-XOR-filter record order / probe arithmetic may differ from "
                     "the ones assumed here (--debug shows them; --ixf-record / --ixf-scheme select others; INTEGRATION.md)." << std::endl;
    for (auto &ch : pool)
        txr_host_free(ch.words);
    for (txr_ctx *c : ctxs)
        txr_ctx_destroy(c);
    if (!parse_error.empty())
        throw std::runtime_error(parse_error);
    if (!worker_error.empty())
        throw std::runtime_error(worker_error);
}

int execute_search(int argc, char const **argv) // taxor_search.cpp:364-389
{
    Config cfg;
    std::vector<std::string> index_files, query_files;
    try
    {
        parse_args(argc, argv, cfg);
        std::cout << "checking input ... " << std::flush;
        // sanity_checks, taxor_search.cpp:97-151
        index_files = str_split(cfg.index_file, ',');
        uint8_t k = 1, s = 1, t = 1;
        uint64_t window = 1;
        uint16_t scaling = 1;
        bool syncmer = false;
        for (auto &f : index_files)
        {
            if (!file_exists(f))
                throw ParserError("Please check the given index file(s). \nThe following index file does not exist: " + f);
            if (index_files.size() > 1)
            {
                TaxorIndexFile idx;
                const std::string e = load_index_file(f, cfg, idx);
                if (!e.empty())
                    throw ParserError(e);
                if (k == 1)
                {
                    k = idx.kmer_size;
                    window = idx.window_size;
                    scaling = idx.scaling;
                    s = idx.syncmer_size;
                    t = idx.t_syncmer;
                    syncmer = idx.use_syncmer;
                    continue;
                }
                if (k != idx.kmer_size || window != idx.window_size || scaling != idx.scaling || s != idx.syncmer_size ||
                    t != idx.t_syncmer || syncmer != idx.use_syncmer)
                    throw ParserError("At least two index files have been created with different kmer selection schemes.\n Please provide only index files using the same kmer-/syncmer-/window-size!");
            }
        }
        query_files = str_split(cfg.query_file, ',');
        for (auto &f : query_files)
            if (!file_exists(f))
                throw ParserError("Please check the given input query files. \nThe following query file does not exist: " + f);
        std::cout << "done!" << std::endl;
    }
    catch (ParserError const &e)
    {
        std::cerr << "[TAXOR SEARCH ERROR] " << e.what() << '\n';
        return -1;
    }

    const std::vector<int> devices = pick_devices(cfg.gpus);
    if (devices.empty())
    {
        std::cerr << "[TAXOR SEARCH ERROR] no CUDA device available: " << txr_last_error() << " (this build has no CPU path)\n";
        return -1;
    }
    // search_hixf, taxor_search.cpp:340-360: one header, then query x index appended
    std::ofstream out(cfg.report_file);
    out << "#QUERY_NAME\tACCESSION\tREFERENCE_NAME\tTAXID\tREF_LEN\tQUERY_LEN\tQHASH_COUNT\tQHASH_MATCH\tTAX_STR\tTAX_ID_STR\n";
    try
    {
        for (auto &q : query_files)
            for (auto &ix : index_files)
                search_single(cfg, q, ix, out, devices);
    }
    catch (std::exception const &e)
    {
        std::cerr << "[TAXOR SEARCH ERROR] " << e.what() << '\n';
        return -1;
    }
    return 0;
}
} // namespace

int main(int argc, char const **argv)
{
    if (argc < 2 || std::string(argv[1]) == "-h" || std::string(argv[1]) == "--help")
    {
        std::cout << "taxor (B200 search driver) 0.2.0\n  usage: taxor search --index-file <a.hixf[,b.hixf]> --query-file <reads.fq[.gz][,...]> "
                     "--output-file <out.tsv> [--threads 1..32] [--percentage 0..1] [--error-rate 0..1]\n"
                     "  `taxor build` and `taxor profile` are provided by the reference binary (CPU).\n";
        return argc < 2 ? -1 : 0;
    }
    const std::string sub = argv[1];
    int rc = 0;
    if (sub == "search")
        rc = execute_search(argc - 2, argv + 2);
    else if (sub == "build" || sub == "profile")
    {
        std::cerr << "[TAXOR ERROR] sub-command '" << sub << "' is not part of the B200 search driver; use the reference taxor binary.\n";
        return -1;
    }
    else
    {
        std::cerr << "[TAXOR ERROR] You either forgot or misspelled the subcommand! Please specify which sub-program you want to use: one of [build, search, profile].\n";
        return -1;
    }
    std::cout << "CPU time  : " << cputime() << " sec" << std::endl;          // main.cpp:79-84
    std::cout << "Peak RSS  : " << (int)(peak_rss() / (1024 * 1024)) << " MByte" << std::endl;
    return rc;
}
