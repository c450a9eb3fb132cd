// kj_cli.cpp -- `kaiju-b200`: the reference's command-line surface (src/kaiju.cpp:74-202, usage 430-451) on top of the C ABI.
//   kaiju-b200 -t nodes.dmp -f db.fmi -i reads.fastq [-j reads2.fastq] [-a mem|greedy] [-m -s -e -E -l] [-x|-X] [-o out] [-z N] [-v]
// Output: "C\t<name>\t<taxid>\n" / "U\t<name>\t0\n" (ConsumerThread.cpp:724-739), in INPUT order.
// Host glue only: option parsing; kj_classify_files() reads FASTA/FASTQ(.gz), parses it on the device with the reference's name
// trimming (kaiju.cpp:318-335) and strip() (util.cpp:26-33), classifies and formats the output.  -z is accepted and ignored (the GPU replaces the consumer threads); -p = protein input; with -v
// columns 4 (best length/score) and 5 (match taxon ids) are appended -- columns 6-7 of the reference's -v output
// (accession names, fragment sequences) are not produced.
#include <getopt.h>
#include <unistd.h>
#include <chrono>
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
                    "   -w FILENAME   Write the device-native index file for -t/-f and exit; such a file can then be given as -f (no -t needed, no transcode at start-up)\n   -T FILENAME   Also write kaiju2table's summary (reads per taxon of rank -r, default species; needs -N names.dmp) from the counts kept on the GPU\n   -p            Input sequences are protein sequences\n   -v            Enable verbose output (adds the match length/score and the matching taxon ids)\n   -d INT        CUDA device ordinal (default 0)\n", prog);
    exit(EXIT_FAILURE);
}

int main(int argc, char** argv) {
    kj_params P; P.mode = 1; P.min_fragment_length = 11; P.mismatches = 3; P.min_score = 65; P.seed_length = 7; P.use_evalue = 1; P.min_evalue = 0.01; P.seg = 1; P.input_is_protein = 0;
    std::string nodes_fn, fmi_fn, in1, in2, out_fn, native_out, table_fn, table_rank = "species", names_fn; bool verbose = false; int device = 0; int c;
    while ((c = getopt(argc, argv, "a:hd:pxXvn:m:e:E:l:t:f:i:j:s:z:o:w:T:r:N:")) != -1) {
        switch (c) {
            case 'a': if (!strcmp(optarg, "mem")) { P.mode = 0; P.use_evalue = 0; } else if (!strcmp(optarg, "greedy")) P.mode = 1; else { fprintf(stderr, "-a must be a valid mode.\n"); usage(argv[0]); } break;
            case 'h': usage(argv[0]); break;
            case 'd': device = atoi(optarg); break;
            case 'v': verbose = true; break;
            case 'p': P.input_is_protein = 1; break;
            case 'x': P.seg = 1; break;
            case 'X': P.seg = 0; break;
            case 'o': out_fn = optarg; break;
            case 'w': native_out = optarg; break;
            case 'T': table_fn = optarg; break;
            case 'r': table_rank = optarg; break;
            case 'N': names_fn = optarg; break;
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
    if (fmi_fn.empty()) { fprintf(stderr, "Error: Please specify the location of the FMI file, using the -f option.\n\n"); usage(argv[0]); }
    // -f may name a device-native index file (written with -w): it holds the taxonomy too, so -t is not needed then
    bool native_in = false;
    { FILE* f = fopen(fmi_fn.c_str(), "rb"); char m[8] = {0}; if (f) { native_in = fread(m, 1, 8, f) == 8 && memcmp(m, "KJB200IX", 8) == 0; fclose(f); } }
    if (nodes_fn.empty() && !native_in) { fprintf(stderr, "Error: Please specify the location of the nodes.dmp file, using the -t option.\n\n"); usage(argv[0]); }
    if (!native_out.empty()) {              // -w FILE: transcode .fmi + nodes.dmp into the device-native index file and exit (no GPU needed)
        if (native_in) die("-w needs the reference's .fmi as -f");
        kj_fmi* fmi = nullptr; kj_nodes* nodes = nullptr;
        if (kj_nodes_load(nodes_fn.c_str(), &nodes) != KJ_OK || kj_fmi_load(fmi_fn.c_str(), &fmi) != KJ_OK) die(kj_last_error());
        kj_index_view iv; kj_taxonomy_view tv; kj_fmi_view(fmi, &iv); kj_nodes_view(nodes, &tv);
        if (kj_native_index_write(&iv, &tv, native_out.c_str()) != KJ_OK) die(kj_last_error());
        kj_fmi_free(fmi); kj_nodes_free(nodes);
        return EXIT_SUCCESS;
    }
    if (in1.empty()) { fprintf(stderr, "Error: Please specify the location of the input file, using the -i option.\n\n"); usage(argv[0]); }
    const bool paired = !in2.empty();
    if (paired && P.input_is_protein) { fprintf(stderr, "Error: Protein input only supports one input file.\n\n"); usage(argv[0]); }      // kaiju.cpp:201

    auto split = [](const std::string& v) { std::vector<std::string> out; size_t b = 0; while (b <= v.size()) { size_t e = v.find(',', b); if (e == std::string::npos) e = v.size(); if (e > b) out.push_back(v.substr(b, e - b)); b = e + 1; } return out; };
    const std::vector<std::string> l1 = split(in1), l2 = split(in2), lo = split(out_fn);
    if (l1.empty()) die("Please specify the location of the input file, using the -i option.");
    if (paired && l2.size() != l1.size()) die("Length of input file lists differ");                      // kaiju-multi.cpp:255-258
    if (!lo.empty() && lo.size() != l1.size()) die("Length of input and output file lists differ");
    if (lo.empty() && l1.size() > 1) die("Several input files need a list of output files (-o)");

    kj_fmi* fmi = nullptr; kj_nodes* nodes = nullptr; kj_ctx* ctx = nullptr;
    if (native_in) { if (kj_create_from_native(&ctx, device, &P, fmi_fn.c_str()) != KJ_OK) die(kj_last_error()); }
    else {
        if (kj_nodes_load(nodes_fn.c_str(), &nodes) != KJ_OK) die(kj_last_error());
        if (kj_fmi_load(fmi_fn.c_str(), &fmi) != KJ_OK) die(kj_last_error());
        kj_index_view iv; kj_taxonomy_view tv; kj_fmi_view(fmi, &iv); kj_nodes_view(nodes, &tv);
        if (kj_create(&ctx, device, &P, &iv, &tv) != KJ_OK) die(kj_last_error());
        kj_fmi_free(fmi); kj_nodes_free(nodes);
    }

    // parsing, classification and output formatting all run on the device; the host moves bytes (kj_ingest.h).
    // Comma-separated lists for -i / -j / -o process several data sets against the index loaded once (kaiju-multi.cpp:220-330).
    if (!table_fn.empty() && (names_fn.empty() || nodes_fn.empty())) die("The summary table (-T) needs names.dmp (-N) and nodes.dmp (-t).");
    for (size_t k = 0; k < l1.size(); k++) {
        uint64_t n_reads = 0, n_classified = 0;
        if (!table_fn.empty() && kj_counts_reset(ctx) != KJ_OK) die(kj_last_error());
        const auto t0 = std::chrono::steady_clock::now();
        if (kj_classify_files(ctx, l1[k].c_str(), paired ? l2[k].c_str() : nullptr, lo.empty() ? nullptr : lo[k].c_str(), verbose ? 1 : 0, &n_reads, &n_classified) != KJ_OK) die(kj_last_error());
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!table_fn.empty()) {     // kaiju2table's report straight from the per-taxon counts in HBM (one block of rows per data set)
            kj_table_opts to; memset(&to, 0, sizeof to); to.rank = table_rank.c_str();
            const std::string label = lo.empty() ? l1[k] : lo[k];
            if (kj_counts_table(ctx, nodes_fn.c_str(), names_fn.c_str(), label.c_str(), &to, table_fn.c_str(), k > 0) != KJ_OK) die(kj_last_error());
        }
        if (verbose || getenv("KJ_CLI_TIMING")) fprintf(stderr, "%s: %llu reads, %llu classified, %.4f s from the first byte read to the last byte written\n", l1[k].c_str(), (unsigned long long)n_reads, (unsigned long long)n_classified, secs);
    }
    kj_destroy(ctx);
    return EXIT_SUCCESS;
}
