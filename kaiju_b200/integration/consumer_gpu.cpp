// consumer_gpu.cpp -- the reference-side binding of INTEGRATION.md sections 1-2 as a compilable translation unit.
//
// It is compiled against the reference's OWN headers (ConsumerThread.hpp, Config.hpp, bwt/bwt.h: -I<reference>/src) and linked with
// the reference's unmodified objects (kaiju.o, Config.o, util.o, ReadItem.o, bwt/*.o) IN PLACE OF ConsumerThread.o, plus
// libkaijub200.so.  The stock main() of kaiju.cpp then runs unchanged: it parses the command line, loads nodes.dmp and the .fmi
// with parseNodesDmp / readFMI, starts `-z` ConsumerThreads and feeds them ReadItems; only what a ConsumerThread DOES with the
// items changes -- it collects them into batches and hands the batches to kj_classify() / kj_classify_verbose()
// (replaces ConsumerThread::doWork, ConsumerThread.cpp:630-749).  The test-infrastructure Makefile builds it as `kaiju-gpu` next to the reference binaries; the GPU test
// tests/test_gpu_binding.py requires its output to equal the stock binary's.  No reference source is copied: the two member
// functions kaiju.cpp references (constructor, doWork) are defined here, the search members of the class are simply not linked.
#include "ConsumerThread.hpp"
#include "kaiju_b200.h"
#include <mutex>

namespace {
std::mutex g_mu;                 // one GPU context for the process; the library context is driven by one thread at a time
kj_ctx* g_ctx = nullptr;

// section 1 of INTEGRATION.md: views straight out of the reference's loader structs (bwt/bwt.h:13-34, compactfmi.h:10-19, suffixArray.h:10-33)
kj_ctx* gpu_context(Config* config) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ctx) return g_ctx;
    BWT* b = config->bwt; FMI* f = config->fmi; suffixArray* s = b->s;
    std::vector<uint64_t> seq_taxon((size_t)b->nseq);
    for (int i = 0; i < b->nseq; i++) {                       // the rule of ConsumerThread.cpp:812-832
        const char* u = strrchr(s->ids[i], '_');
        unsigned long v = strtoul(u ? u + 1 : s->ids[i], NULL, 10);
        seq_taxon[(size_t)i] = (v == ULONG_MAX) ? UINT64_MAX : (uint64_t)v;
    }
    kj_index_view iv; memset(&iv, 0, sizeof iv);
    iv.alen = f->alen; iv.alphabet = b->alphabet; iv.bwtlen = f->bwtlen; iv.bwt = f->bwt; iv.startLcode = f->startLcode;
    iv.db_len = b->len; iv.nseq = b->nseq; iv.ncheck = s->ncheck; iv.chpt_exp = s->chpt_exp; iv.nbytes = s->nbytes; iv.pbits = s->pbits;
    iv.sa = s->sa; iv.seq_taxon = seq_taxon.data();
    std::vector<uint64_t> node, parent;
    for (auto& kv : *config->nodes) { node.push_back(kv.first); parent.push_back(kv.second); }
    kj_taxonomy_view tv; tv.n = node.size(); tv.node = node.data(); tv.parent = parent.data();
    kj_params p; memset(&p, 0, sizeof p);
    p.mode = config->mode == MEM ? 0 : 1; p.min_fragment_length = config->min_fragment_length; p.mismatches = config->mismatches;
    p.min_score = config->min_score; p.seed_length = config->seed_length; p.use_evalue = config->use_Evalue ? 1 : 0; p.min_evalue = config->min_Evalue;
    p.seg = config->SEG ? 1 : 0; p.input_is_protein = config->input_is_protein ? 1 : 0;
    const char* dev = getenv("KAIJU_GPU_DEVICE");
    if (kj_create(&g_ctx, dev ? atoi(dev) : 0, &p, &iv, &tv) != KJ_OK) { error(kj_last_error()); exit(EXIT_FAILURE); }
    return g_ctx;
}
struct Batch {
    std::string seq1, seq2; std::vector<uint64_t> off1{0}, off2{0}; std::vector<std::string> names;
    void add(const ReadItem* it) { names.push_back(it->name); seq1 += it->sequence1; off1.push_back(seq1.size()); seq2 += it->sequence2; off2.push_back(seq2.size()); }
    void clear() { seq1.clear(); seq2.clear(); off1.assign(1, 0); off2.assign(1, 0); names.clear(); }
};
}  // namespace

ConsumerThread::ConsumerThread(ProducerConsumerQueue<ReadItem*>* workQueue, Config* config) : myWorkQueue(workQueue), config(config) {}

// section 2 of INTEGRATION.md: ReadItems -> batch -> kj_classify -> the output lines of ConsumerThread.cpp:724-739
void ConsumerThread::doWork() {
    kj_ctx* ctx = gpu_context(config);
    Batch b; bool paired = false; ReadItem* item = NULL;
    auto flush = [&]() {
        const size_t n = b.names.size(); if (!n) return;
        std::vector<uint64_t> taxon(n); std::vector<uint32_t> best(n); std::vector<uint64_t> ids; std::vector<uint8_t> nids;
        int rc;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            if (config->verbose) {
                ids.resize(n * KJ_MAX_MATCH_IDS); nids.resize(n);
                rc = kj_classify_verbose(ctx, b.seq1.data(), b.off1.data(), paired ? b.seq2.data() : nullptr, paired ? b.off2.data() : nullptr, n, taxon.data(), best.data(), ids.data(), nids.data());
            } else rc = kj_classify(ctx, b.seq1.data(), b.off1.data(), paired ? b.seq2.data() : nullptr, paired ? b.off2.data() : nullptr, n, taxon.data(), nullptr);
        }
        if (rc != KJ_OK) { error(kj_last_error()); exit(EXIT_FAILURE); }
        for (size_t i = 0; i < n; i++) {
            if (!taxon[i]) { output << "U\t" << b.names[i] << "\t0\n"; continue; }
            output << "C\t" << b.names[i] << "\t" << taxon[i];
            if (config->verbose) {          // columns 4-5 of `kaiju -v` (ConsumerThread.cpp:527-536, 614-623); the accession / fragment columns are not produced by the library
                output << "\t" << best[i] << "\t";
                for (uint8_t k = 0; k < nids[i]; k++) output << ids[i * KJ_MAX_MATCH_IDS + k] << ",";
            }
            output << "\n";
        }
        {   // flush_output (ConsumerThread.cpp:847-856)
            static std::mutex m; std::lock_guard<std::mutex> out_lock(m);
            *(config->out_stream) << output.str();
        }
        output.str(""); b.clear();
    };
    while (myWorkQueue->pop(&item)) {
        paired = paired || item->paired;
        b.add(item); delete item;
        if (b.names.size() >= (1u << 18)) flush();
    }
    flush();
}
