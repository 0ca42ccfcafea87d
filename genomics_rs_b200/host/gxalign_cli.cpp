// gxalign_cli.cpp -- C++ host above the C ABI: the `align` sub-command of the reference CLI
// (/root/reference/src/main.rs:27-84,115-153) on top of libgxalign.
//
//   gxalign_cli [--config-path config.toml] align [-a local|1|<anything else = global>] -f pair.fasta
//   gxalign_cli [--config-path config.toml] align-all [-a ...] --fasta-dir DIR      
//
// align-all ingests a directory the way the reference's `compare` sub-command does (main.rs:227-239: every
// *.fasta file, all of its records) -- in SORTED file-name order, because NW with the reference's tie-breaks is not
// symmetric and read_dir order is unspecified -- and aligns every pair (a < b, s1 = a) in one gx_align_batch call.
//
// Host-side restatements: FASTA loader (src/sequence.rs:45-95), config (src/config.rs:21-40, TOML subset:
// one [scores] table with four integer keys), Display for AlignedSequences (src/alignment/display.rs:9-127).
// All alignment work happens in gx_align_pair; without an sm_100 GPU the program exits with status 1.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/gxalign.h"

struct Sequence { std::string name, sequence; };

static std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

// sequence.rs:45-95
static std::vector<Sequence> from_fasta(const std::string &path) {
    std::vector<Sequence> seqs;
    std::ifstream in(path, std::ios::binary);
    if (!in) {
        fprintf(stderr, "ERROR Could not open file: %s\n", path.c_str());
        return seqs;
    }
    std::string line;
    bool have = false;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        if (line[0] == '>') {
            seqs.push_back({trim(line.substr(1)), ""});
            have = true;
        } else if (have) {
            seqs.back().sequence += trim(line);
        } else {
            fprintf(stderr, "WARN Sequence data found without a header\n");
        }
    }
    return seqs;
}

// config.rs:21-40; exit(1) on read or parse failure
static gx_scores get_config(const std::string &path) {
    std::ifstream in(path);
    if (!in) {
        fprintf(stderr, "ERROR Could not read config file: %s\n", path.c_str());
        exit(1);
    }
    gx_scores sc{};
    bool in_scores = false;
    int seen = 0;
    std::string line;
    while (std::getline(in, line)) {
        std::string t = trim(line.substr(0, line.find('#')));
        if (t.empty()) continue;
        if (t.front() == '[') {
            in_scores = (t == "[scores]");
            continue;
        }
        size_t eq = t.find('=');
        if (!in_scores || eq == std::string::npos) continue;
        std::string key = trim(t.substr(0, eq)), val = trim(t.substr(eq + 1));
        char *end = nullptr;
        long v = strtol(val.c_str(), &end, 10);
        if (end == val.c_str() || *end) {
            fprintf(stderr, "ERROR Could not parse config file: %s\n", path.c_str());
            exit(1);
        }
        if (key == "s_match") sc.s_match = (int32_t)v, seen |= 1;
        else if (key == "s_mismatch") sc.s_mismatch = (int32_t)v, seen |= 2;
        else if (key == "g") sc.g = (int32_t)v, seen |= 4;
        else if (key == "h") sc.h = (int32_t)v, seen |= 8;
    }
    if (seen != 15) {
        fprintf(stderr, "ERROR Could not parse config file: %s\n", path.c_str());
        exit(1);
    }
    return sc;
}

static std::string rust_f64(double x) {  // Rust `{}`: shortest round-trip, no exponent, no trailing ".0"
    if (std::isnan(x)) return "NaN";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}
static std::string pct2(uint64_t num, uint64_t den) {
    if (!den) return "NaN";
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", (double)num / (double)den * 100.0);
    return buf;
}

// display.rs:9-127
static std::string display(const Sequence &s1, const Sequence &s2, const gx_result &r, const std::vector<uint8_t> &ops) {
    const size_t W = 200;
    std::ostringstream f;
    std::string a, mid, b;
    size_t i1 = 0, i2 = 0, hl = 0, idx = 0, n = ops.size();
    static const char sym[6] = {'|', 'x', ' ', ' ', '%', '%'};
    while (idx < n) {
        uint8_t c = ops[n - 1 - idx];
        if (hl > W) {
            f << "\n\n" << idx - W << "-" << idx << ":\n\n" << a << "\n" << mid << "\n" << b << "\n";
            a.clear(); mid.clear(); b.clear();
            hl = 0;
        }
        if (c == GX_INSERT || c == GX_OPEN_INSERT) a.push_back('-');
        else if (i1 < s1.sequence.size()) a.push_back(s1.sequence[i1++]);
        mid.push_back(sym[c]);
        if (c == GX_DELETE || c == GX_OPEN_DELETE) b.push_back('-');
        else if (i2 < s2.sequence.size()) b.push_back(s2.sequence[i2++]);
        hl++; idx++;
    }
    f << "\n\n" << idx - a.size() << "-" << idx << ":\n\n" << a << "\n" << mid << "\n" << b << "\n";
    f << "\n\nAlignment Score: " << r.score << "\n";
    f << "Matches: " << r.matches << "/" << idx << " (" << pct2(r.matches, idx) << "%)\n";
    f << "Mismatches: " << r.mismatches << "/" << idx << " (" << pct2(r.mismatches, idx) << "%)\n";
    f << "Gap Extensions: " << r.gap_extensions << "/" << idx << " (" << pct2(r.gap_extensions, idx) << "%)\n";
    f << "Opening Gaps: " << r.opening_gaps << "/" << idx << " (" << pct2(r.opening_gaps, idx) << "%)\n";
    f << "Percent Identity " << rust_f64(idx ? (double)r.matches / (double)idx * 100.0 : NAN) << "%\n";
    return f.str();
}

// every *.fasta file of `dir`, sorted by file name, all records of each (main.rs:230-237)
static std::vector<Sequence> from_fasta_dir(const std::string &dir) {
    std::vector<std::string> files;
    if (DIR *d = opendir(dir.c_str())) {
        while (dirent *e = readdir(d)) {
            const std::string name = e->d_name;
            if (name.size() > 6 && name.compare(name.size() - 6, 6, ".fasta") == 0) files.push_back(name);
        }
        closedir(d);
    } else {
        fprintf(stderr, "ERROR Could not open directory: %s\n", dir.c_str());
    }
    std::sort(files.begin(), files.end());
    std::vector<Sequence> all;
    for (const auto &f : files) {
        std::vector<Sequence> s = from_fasta(dir + "/" + f);
        all.insert(all.end(), s.begin(), s.end());
    }
    return all;
}

static int align_all(const std::vector<Sequence> &seqs, gx_scores sc, bool is_local) {
    const size_t ns = seqs.size();
    if (ns < 2) {
        fprintf(stderr, "ERROR need at least two sequences\n");
        return 101;
    }
    std::vector<uint8_t> blob;
    std::vector<uint64_t> off(ns), len(ns);
    for (size_t k = 0; k < ns; ++k) {
        off[k] = blob.size();
        len[k] = seqs[k].sequence.size();
        blob.insert(blob.end(), seqs[k].sequence.begin(), seqs[k].sequence.end());
    }
    std::vector<uint64_t> off1, len1, off2, len2, ops_off(1, 0);
    std::vector<std::pair<size_t, size_t>> jobs;
    for (size_t a = 0; a < ns; ++a)
        for (size_t b = a + 1; b < ns; ++b) {
            jobs.push_back({a, b});
            off1.push_back(off[a]); len1.push_back(len[a]);
            off2.push_back(off[b]); len2.push_back(len[b]);
            ops_off.push_back(ops_off.back() + len[a] + len[b] + 1);
        }
    std::vector<gx_result> res(jobs.size());
    std::vector<uint8_t> ops(ops_off.back());
    int rc = gx_align_batch(blob.data(), blob.size(), off1.data(), len1.data(), off2.data(), len2.data(), jobs.size(), sc, is_local,
                            GX_FLAG_TRACEBACK, res.data(), ops.data(), ops_off.data());
    if (rc) {
        fprintf(stderr, "ERROR %s: %s\n", gx_strerror(rc), gx_last_error());
        return 1;
    }
    printf("# %zu sequences, %zu pairs, %s; fill %.3f ms + walk %.3f ms on the GPU\n", ns, jobs.size(), is_local ? "local" : "global",
           res[0].fill_ms, res[0].walk_ms);
    printf("s1\ts2\tscore\tops\tmatches\tmismatches\tgap_extensions\topening_gaps\tidentity\n");
    for (size_t q = 0; q < jobs.size(); ++q) {
        const gx_result &r = res[q];
        printf("%s\t%s\t%lld\t%llu\t%llu\t%llu\t%llu\t%llu\t%.4f\n", seqs[jobs[q].first].name.c_str(), seqs[jobs[q].second].name.c_str(),
               (long long)r.score, (unsigned long long)r.n_ops, (unsigned long long)r.matches, (unsigned long long)r.mismatches,
               (unsigned long long)r.gap_extensions, (unsigned long long)r.opening_gaps, r.n_ops ? (double)r.matches / (double)r.n_ops : 0.0);
    }
    return 0;
}

int main(int argc, char **argv) {
    std::string config_path = "config.toml", type = "local", fasta, fasta_dir;   // main.rs:31-32: default local
    bool align = false, all = false;
    for (int k = 1; k < argc; ++k) {
        std::string a = argv[k];
        if ((a == "--config-path" || a == "-c") && k + 1 < argc) config_path = argv[++k];
        else if (a == "align") align = true;
        else if (a == "align-all") all = true;
        else if ((a == "--fasta-dir" || a == "-d") && k + 1 < argc) fasta_dir = argv[++k];
        else if ((a == "-a" || a == "--alignment-type") && k + 1 < argc) type = argv[++k];
        else if ((a == "-f" || a == "--fasta-path") && k + 1 < argc) fasta = argv[++k];
    }
    if (!((align && !fasta.empty()) || (all && !fasta_dir.empty()))) {
        fprintf(stderr, "usage: %s [--config-path config.toml] align [-a local|1|global] -f pair.fasta\n"
                        "       %s [--config-path config.toml] align-all [-a local|1|global] --fasta-dir DIR\n", argv[0], argv[0]);
        return 2;
    }
    gx_scores sc = get_config(config_path);
    if (all) {
        int rc0 = gx_init(-1);
        if (rc0) {
            fprintf(stderr, "ERROR %s: %s\n", gx_strerror(rc0), gx_last_error());
            return 1;
        }
        const int rca = align_all(from_fasta_dir(fasta_dir), sc, type == "local" || type == "1");
        gx_shutdown();
        return rca;
    }
    std::vector<Sequence> seqs = from_fasta(fasta);
    if (seqs.size() > 2) fprintf(stderr, "WARN More than two sequences found. Only the first two will be used.\n");
    if (seqs.size() < 2) {
        fprintf(stderr, "ERROR need two sequences (the reference panics here: algo.rs:168-169)\n");
        return 101;
    }
    const bool is_local = (type == "local" || type == "1");   // main.rs:142
    int rc = gx_init(-1);
    if (rc) {
        fprintf(stderr, "ERROR %s: %s\n", gx_strerror(rc), gx_last_error());
        return 1;
    }
    const std::string &s1 = seqs[0].sequence, &s2 = seqs[1].sequence;
    std::vector<uint8_t> ops(s1.size() + s2.size() + 1);
    gx_result r;
    rc = gx_align_pair((const uint8_t *)s1.data(), s1.size(), (const uint8_t *)s2.data(), s2.size(), sc, is_local,
                       GX_FLAG_TRACEBACK, &r, ops.data(), ops.size());
    if (rc) {
        fprintf(stderr, "ERROR %s: %s\n", gx_strerror(rc), gx_last_error());
        return 1;
    }
    ops.resize(r.n_ops);
    fprintf(stderr, "INFO Sequence table shape: [%zu, %zu]\n", s1.size() + 1, s2.size() + 1);
    fprintf(stderr, "INFO Table initialization complete, time taken: %lldus\n", (long long)(r.fill_ms * 1000));
    fprintf(stderr, "INFO Starting at (%llu, %llu)\n", (unsigned long long)r.start_i, (unsigned long long)r.start_j);
    fprintf(stderr, "INFO Retrace complete, time taken: %lldus\n", (long long)(r.walk_ms * 1000));
    fputs(display(seqs[0], seqs[1], r, ops).c_str(), stdout);
    gx_shutdown();
    return 0;
}
