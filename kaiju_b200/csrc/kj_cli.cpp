// kj_cli.cpp -- `kaiju-b200`: the reference's command-line surface (src/kaiju.cpp:74-202, usage 430-451) on top of the C ABI.
//   kaiju-b200 -t nodes.dmp -f db.fmi -i reads.fastq [-j reads2.fastq] [-a mem|greedy] [-m -s -e -E -l] [-x|-X] [-o out] [-z N] [-v]
// Output: "C\t<name>\t<taxid>\n" / "U\t<name>\t0\n" (ConsumerThread.cpp:724-739), in INPUT order.
// Host glue only: FASTA/FASTQ(.gz) parsing with the reference's name trimming (kaiju.cpp:318-335) and strip() (util.cpp:26-33),
// batching, kj_classify().  -z is accepted and ignored (the GPU replaces the consumer threads); -p = protein input; with -v
// columns 4 (best length/score) and 5 (match taxon ids) are appended -- columns 6-7 of the reference's -v output
// (accession names, fragment sequences) are not produced.
#include <getopt.h>
#include <unistd.h>
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "kaiju_b200.h"

static void die(const std::string& m) { fprintf(stderr, "Error: %s\n\n", m.c_str()); exit(EXIT_FAILURE); }
static void usage(const char* prog) {
    fprintf(stderr, "kaiju-b200 (B200-native classification path of Kaiju)\n\nUsage:\n   %s -t nodes.dmp -f kaiju_db.fmi -i reads.fastq [-j reads2.fastq]\n\n"
                    "Mandatory arguments:\n   -t FILENAME   Name of nodes.dmp file\n   -f FILENAME   Name of database (.fmi) file\n   -i FILENAME   Name of input file containing reads in FASTA or FASTQ format\n\n"
                    "Optional arguments:\n   -j FILENAME   Name of second input file for paired-end reads\n   -o FILENAME   Name of output file. If not specified, output will be printed to STDOUT\n"
                    "   -z INT        accepted for compatibility (ignored: the GPU replaces the worker threads)\n   -a STRING     Run mode, either \"mem\"  or \"greedy\" (default: greedy)\n"
                    "   -e INT        Number of mismatches allowed in Greedy mode (default: 3)\n   -m INT        Minimum match length (default: 11)\n   -s INT        Minimum match score in Greedy mode (default: 65)\n"
                    "   -E FLOAT      Minimum E-value in Greedy mode (default: 0.01)\n   -x            Enable SEG low complexity filter (enabled by default)\n   -X            Disable SEG low complexity filter\n"
                    "   -p            Input sequences are protein sequences\n   -v            Enable verbose output (adds the match length/score and the matching taxon ids)\n   -d INT        CUDA device ordinal (default 0)\n", prog);
    exit(EXIT_FAILURE);
}

struct Reader {
    gzFile fp = nullptr; std::string path; bool first = true, fastq = false; std::string line, pending; bool has_pending = false; std::vector<char> buf;
    explicit Reader(const std::string& p) : path(p), buf(1 << 16) { fp = gzopen(p.c_str(), "rb"); if (!fp) die("Could not open file " + p); gzbuffer(fp, 1 << 20); }
    ~Reader() { if (fp) gzclose(fp); }
    bool getline(std::string& out) {
        if (has_pending) { out.swap(pending); has_pending = false; return true; }
        out.clear();
        for (;;) {
            if (!gzgets(fp, buf.data(), (int)buf.size())) return !out.empty();
            size_t l = strlen(buf.data()); out.append(buf.data(), l);
            if (l && out.back() == '\n') { out.pop_back(); return true; }
            if (l + 1 < buf.size()) return true;          // EOF without newline
        }
    }
    void unget(std::string& l) { pending.swap(l); has_pending = true; }
    // next record: name (trimmed at " /\t\r") + sequence stripped of non-letters; false at EOF
    bool next(std::string& name, std::string& seq) {
        do { if (!getline(line)) return false; } while (line.empty());
        if (first) { if (line[0] == '@') fastq = true; else if (line[0] != '>') die("Auto-detection of file type for file " + path + " failed."); first = false; }
        line.erase(0, 1); size_t n = line.find_first_of(" /\t\r"); if (n != std::string::npos) line.erase(n);
        name = line; seq.clear();
        if (fastq) { std::string s, skip; getline(s); getline(skip); getline(skip); append_stripped(seq, s); }
        else { std::string s; while (getline(s)) { if (!s.empty() && s[0] == '>') { unget(s); break; } append_stripped(seq, s); } }
        return true;
    }
    static void append_stripped(std::string& dst, const std::string& s) { for (char c : s) if ((c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z')) dst.push_back(c); }
};

int main(int argc, char** argv) {
    kj_params P; P.mode = 1; P.min_fragment_length = 11; P.mismatches = 3; P.min_score = 65; P.seed_length = 7; P.use_evalue = 1; P.min_evalue = 0.01; P.seg = 1; P.input_is_protein = 0;
    std::string nodes_fn, fmi_fn, in1, in2, out_fn; bool verbose = false; int device = 0; int c;
    while ((c = getopt(argc, argv, "a:hd:pxXvn:m:e:E:l:t:f:i:j:s:z:o:")) != -1) {
        switch (c) {
            case 'a': if (!strcmp(optarg, "mem")) { P.mode = 0; P.use_evalue = 0; } else if (!strcmp(optarg, "greedy")) P.mode = 1; else { fprintf(stderr, "-a must be a valid mode.\n"); usage(argv[0]); } break;
            case 'h': usage(argv[0]); break;
            case 'd': device = atoi(optarg); break;
            case 'v': verbose = true; break;
            case 'p': P.input_is_protein = 1; break;
            case 'x': P.seg = 1; break;
            case 'X': P.seg = 0; break;
            case 'o': out_fn = optarg; break;
            case 'f': fmi_fn = optarg; break;
            case 't': nodes_fn = optarg; break;
            case 'i': in1 = optarg; break;
            case 'j': in2 = optarg; break;
            case 'l': { int v = atoi(optarg); if (v < 7) { die("Seed length must be >= 7."); } P.seed_length = (uint32_t)v; break; }
            case 's': { int v = atoi(optarg); if (v <= 0) die("Min Score (-s) must be greater than 0."); P.min_score = (uint32_t)v; break; }
            case 'm': { int v = atoi(optarg); if (v <= 0) die("Min fragment length (-m) must be greater than 0."); P.min_fragment_length = (uint32_t)v; break; }
            case 'e': { int v = atoi(optarg); if (v < 0) die("Number of mismatches must be >= 0."); P.mismatches = (uint32_t)v; break; }
            case 'E': { P.min_evalue = atof(optarg); if (P.min_evalue <= 0.0) die("E-value threshold must be greater than 0."); break; }
            case 'z': { if (atoi(optarg) <= 0) die("Number of threads (-z) must be greater than 0."); break; }
            case 'n': break;
            default: usage(argv[0]);
        }
    }
    if (nodes_fn.empty()) { fprintf(stderr, "Error: Please specify the location of the nodes.dmp file, using the -t option.\n\n"); usage(argv[0]); }
    if (fmi_fn.empty()) { fprintf(stderr, "Error: Please specify the location of the FMI file, using the -f option.\n\n"); usage(argv[0]); }
    if (in1.empty()) { fprintf(stderr, "Error: Please specify the location of the input file, using the -i option.\n\n"); usage(argv[0]); }
    const bool paired = !in2.empty();
    if (paired && P.input_is_protein) { fprintf(stderr, "Error: Protein input only supports one input file.\n\n"); usage(argv[0]); }      // kaiju.cpp:201

    kj_fmi* fmi = nullptr; kj_nodes* nodes = nullptr; kj_ctx* ctx = nullptr;
    if (kj_nodes_load(nodes_fn.c_str(), &nodes) != KJ_OK) die(kj_last_error());
    if (kj_fmi_load(fmi_fn.c_str(), &fmi) != KJ_OK) die(kj_last_error());
    kj_index_view iv; kj_taxonomy_view tv; kj_fmi_view(fmi, &iv); kj_nodes_view(nodes, &tv);
    if (kj_create(&ctx, device, &P, &iv, &tv) != KJ_OK) die(kj_last_error());
    kj_fmi_free(fmi); kj_nodes_free(nodes);

    FILE* out = stdout;
    if (!out_fn.empty()) { out = fopen(out_fn.c_str(), "w"); if (!out) die("Could not open file " + out_fn + " for writing"); }
    setvbuf(out, nullptr, _IOFBF, 1 << 22);
    Reader r1(in1); Reader* r2 = paired ? new Reader(in2) : nullptr;
    const size_t BATCH = 1u << 20;
    std::string seq1, seq2, name, name2, s; std::vector<uint64_t> off1, off2, taxon, ids; std::vector<uint32_t> best; std::vector<uint8_t> nids; std::vector<std::string> names;
    auto flush = [&]() {
        size_t n = names.size(); if (!n) return;
        taxon.resize(n); best.resize(n);
        int rc;
        if (verbose) { ids.resize(n * KJ_MAX_MATCH_IDS); nids.resize(n);
                       rc = kj_classify_verbose(ctx, seq1.data(), off1.data(), paired ? seq2.data() : nullptr, paired ? off2.data() : nullptr, n, taxon.data(), best.data(), ids.data(), nids.data()); }
        else rc = kj_classify(ctx, seq1.data(), off1.data(), paired ? seq2.data() : nullptr, paired ? off2.data() : nullptr, n, taxon.data(), best.data());
        if (rc != KJ_OK) die(kj_last_error());
        for (size_t i = 0; i < n; i++) {
            if (taxon[i]) {
                fprintf(out, "C\t%s\t%llu", names[i].c_str(), (unsigned long long)taxon[i]);
                if (verbose) { fprintf(out, "\t%u\t", best[i]); for (unsigned k = 0; k < nids[i]; k++) fprintf(out, "%llu,", (unsigned long long)ids[i * KJ_MAX_MATCH_IDS + k]); }     // ConsumerThread.cpp:527-536
                fputc('\n', out);
            }
            else fprintf(out, "U\t%s\t0\n", names[i].c_str());
        }
        seq1.clear(); seq2.clear(); off1.assign(1, 0); off2.assign(1, 0); names.clear();
    };
    off1.assign(1, 0); off2.assign(1, 0);
    while (r1.next(name, s)) {
        seq1 += s; off1.push_back(seq1.size());
        if (paired) {
            if (!r2->next(name2, s)) die("File " + in1 + " contains more reads then file " + in2);
            if (name != name2) die("Read names are not identical between the two input files. Probably reads are not in the same order in both files.");
            seq2 += s; off2.push_back(seq2.size());
        }
        names.push_back(name);
        if (names.size() >= BATCH || seq1.size() >= (1u << 30)) flush();
    }
    flush();
    if (paired && r2->next(name2, s)) fprintf(stderr, "Warning: File %s has more reads then file %s\n", in2.c_str(), in1.c_str());
    if (out != stdout) fclose(out); else fflush(out);
    delete r2; kj_destroy(ctx);
    return EXIT_SUCCESS;
}
