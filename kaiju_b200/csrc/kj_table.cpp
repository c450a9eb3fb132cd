// kj_table.cpp -- kaiju2table's summary report (src/kaiju2table.cpp:150-365) written from per-taxon read counts
// (the vector a context accumulates in HBM, kj_counts_get) instead of re-reading the per-read output file.
// Host code: the count vector is tiny; what is replaced is the reference's pass over every output line.
// Own implementation of the observable rules:
//   * nodes.dmp with ranks / names.dmp ("scientific name" lines) parsed with the reference's tolerant rules (util.cpp:123-178);
//   * reads of taxa missing from nodes.dmp are warned about and only count towards the total (kaiju2table.cpp:197-200);
//   * counts are rolled up to all ancestors except below Viruses (taxon 10239), which are listed on their own (218-229);
//   * rows = taxa of the requested rank, by descending count (ties: ascending taxon id), filtered by -m / -c; then the
//     Viruses / "cannot be assigned" / threshold / unclassified rows, with the reference's float arithmetic for the percentages.
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <list>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>
#include "kj_host.h"

namespace {
const uint64_t kViruses = 10239;
typedef std::unordered_map<uint64_t, uint64_t> NodeMap;

bool parse_uint(const std::string& s, size_t b, size_t e, uint64_t& v) {
    if (b == std::string::npos || b >= s.size()) return false;
    if (e == std::string::npos) e = s.size();
    if (e <= b) return false;
    v = 0; for (size_t i = b; i < e; i++) { if (s[i] < '0' || s[i] > '9') return false; v = v * 10 + (uint64_t)(s[i] - '0'); }
    return true;
}
int load_nodes(const char* path, NodeMap& nodes, std::unordered_map<uint64_t, std::string>& rank) {
    std::ifstream f(path); if (!f.is_open()) { kj_err() = std::string("Could not open file ") + path; return KJ_ERR_IO; }
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        size_t end = line.find_first_not_of("0123456789"); uint64_t node, parent;
        if (!parse_uint(line, 0, end, node)) continue;
        size_t start = line.find_first_of("0123456789", end); if (start == std::string::npos) continue;
        end = line.find_first_not_of("0123456789", start + 1);
        if (!parse_uint(line, start, end, parent)) continue;
        start = line.find_first_of("abcdefghijklmnopqrstuvwxyz", end);
        if (start == std::string::npos) continue;                  // the reference's substr() throws here and the line is dropped
        size_t e2 = line.find_first_not_of("abcdefghijklmnopqrstuvwxyz ", start);
        nodes.emplace(node, parent); rank.emplace(node, line.substr(start, e2 == std::string::npos ? std::string::npos : e2 - start));
    }
    return KJ_OK;
}
int load_names(const char* path, std::unordered_map<uint64_t, std::string>& names) {
    std::ifstream f(path); if (!f.is_open()) { kj_err() = std::string("Could not open file ") + path; return KJ_ERR_IO; }
    std::string line;
    while (std::getline(f, line)) {
        if (line.empty() || line.find("scientific name") == std::string::npos) continue;
        size_t start = line.find_first_of("0123456789"); if (start == std::string::npos) continue;
        size_t end = line.find_first_not_of("0123456789", start); uint64_t id;
        if (!parse_uint(line, start, end, id)) continue;
        start = line.find_first_not_of("\t|", end); if (start == std::string::npos) continue;
        end = line.find_first_of("\t|", start + 1);
        names.emplace(id, line.substr(start, end == std::string::npos ? std::string::npos : end - start));
    }
    return KJ_OK;
}
// node1 is node2 or one of its ancestors (util.cpp:63-77); false (with the reference's message) if either is unknown
// (the reference prints its "not found in taxonomy" message on every call; here it is said once per report)
bool is_ancestor(const NodeMap& nodes, uint64_t node1, uint64_t node2, bool* warned) {
    if (!nodes.count(node1) || !nodes.count(node2)) {
        if (warned && !*warned) { fprintf(stderr, "Taxon ID %" PRIu64 " not found in taxonomy!\n", nodes.count(node1) ? node2 : node1); *warned = true; }
        return false;
    }
    if (node1 == node2) return true;
    for (auto it = nodes.find(node2); it != nodes.end() && it->second != node2; it = nodes.find(node2)) { node2 = it->second; if (node2 == node1) return true; }
    return false;
}
std::string taxon_name(const std::unordered_map<uint64_t, std::string>& names, uint64_t id, const char* names_path) {
    auto it = names.find(id);
    if (it != names.end()) return it->second;
    fprintf(stderr, "Warning: Taxon ID %" PRIu64 " is not found in file %s.\n", id, names_path);
    return "taxonid:" + std::to_string(id);
}
}  // namespace

extern "C" int kj_table_write(const uint64_t* taxon_ids, const uint64_t* counts, uint64_t n, const char* nodes_path, const char* names_path,
                              const char* label, const kj_table_opts* o, const char* out_path, int append) {
    if ((!taxon_ids || !counts) && n) { kj_err() = "kj_table_write: null argument"; return KJ_ERR_ARG; }
    if (!nodes_path || !names_path || !label || !o || !out_path || !o->rank) { kj_err() = "kj_table_write: null argument"; return KJ_ERR_ARG; }
    const std::string rank = o->rank; const float min_percent = (float)o->min_percent; const int min_read_count = o->min_read_count;
    static const char* kRanks[] = {"phylum", "class", "order", "family", "genus", "species"};
    bool okr = false; for (const char* r : kRanks) okr = okr || rank == r;
    if (!okr) { kj_err() = "Rank must be one of: phylum, class, order, family, genus, species."; return KJ_ERR_ARG; }
    if (min_read_count < 0) { kj_err() = "Min required read count (-c) must be >= 0"; return KJ_ERR_ARG; }
    if (min_percent < 0.0f || min_percent > 100.0f) { kj_err() = "Min required percent (-m) must be between 0.0 and 100.0"; return KJ_ERR_ARG; }
    if (min_percent > 0.0f && min_read_count > 0) { kj_err() = "Either specify minimum percent with -m or minimum read count with -c."; return KJ_ERR_ARG; }
    const bool specified_ranks = o->rank_list && *o->rank_list;
    if (specified_ranks && o->full_path) { kj_err() = "Please use either option -p or -l, but not both of them."; return KJ_ERR_ARG; }
    std::list<std::string> ranks_list; std::set<std::string> ranks_set;
    if (specified_ranks) {
        const std::string a = o->rank_list; size_t b = 0;
        while (b <= a.size()) { size_t e = a.find(',', b); if (e == std::string::npos) e = a.size(); if (e > b) { ranks_list.push_back(a.substr(b, e - b)); ranks_set.insert(a.substr(b, e - b)); } b = e + 1; }
        if (!ranks_set.count(rank)) { kj_err() = "Specified rank " + rank + " is not contained in rank list supplied with option -l"; return KJ_ERR_ARG; }
    }
    NodeMap nodes; std::unordered_map<uint64_t, std::string> node2rank, node2name;
    int rc = load_nodes(nodes_path, nodes, node2rank); if (rc) return rc;
    rc = load_names(names_path, node2name); if (rc) return rc;

    // per-read pass of the reference (186-214), on counts
    bool warned = false;
    std::map<uint64_t, uint64_t> hits; uint64_t unclassified = 0, totalreads = 0, total_virus_reads = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t id = taxon_ids[i], c = counts[i]; if (!c) continue;
        totalreads += c;
        if (id == 0) { unclassified += c; continue; }
        if (!nodes.count(id)) { fprintf(stderr, "Warning: Taxon ID %" PRIu64 " is not contained in %s.\n", id, nodes_path); continue; }
        if (is_ancestor(nodes, kViruses, id, &warned)) total_virus_reads += c;
        hits[id] += c;
    }
    // ancestors get the reads of their descendants, except below Viruses (216-229)
    std::map<uint64_t, uint64_t> summarized;
    for (const auto& h : hits) {
        uint64_t id = h.first; const uint64_t reads = h.second;
        if (is_ancestor(nodes, kViruses, id, &warned)) { summarized[id] = reads; continue; }
        for (auto it = nodes.find(id); it != nodes.end() && it->second != id; it = nodes.find(id)) { summarized[id] += reads; id = it->second; }
    }
    if (o->filter_unclassified) totalreads -= unclassified;
    uint64_t at_rank = 0, below_percent = 0, below_count = 0;
    std::multimap<uint64_t, uint64_t, std::greater<uint64_t>> sorted;
    for (const auto& s : summarized) {
        const uint64_t id = s.first, count = s.second;
        if (is_ancestor(nodes, kViruses, id, &warned)) { sorted.emplace(count, id); continue; }       // viruses are always listed
        auto rk = node2rank.find(id);
        if (rk == node2rank.end()) { fprintf(stderr, "Error: No rank specified for taxonid %" PRIu64 "\n", id); continue; }
        if (rank == rk->second) {
            if ((int)count >= min_read_count) {
                const float percent = (float)count / (float)totalreads * 100;
                if (percent >= min_percent) sorted.emplace(count, id); else below_percent += count;
            } else below_count += count;
            at_rank += count;
        }
    }
    uint64_t above = o->filter_unclassified ? totalreads - at_rank : totalreads - unclassified - at_rank;
    above -= total_virus_reads;

    FILE* f = fopen(out_path, append ? "a" : "w");
    if (!f) { kj_err() = std::string("Could not open file ") + out_path + " for writing"; return KJ_ERR_IO; }
    if (!append) fprintf(f, "file\tpercent\treads\ttaxon_id\ttaxon_name\n");
    for (const auto& e : sorted) {
        if (!o->expand_viruses && is_ancestor(nodes, kViruses, e.second, &warned)) continue;
        const float percent = (float)e.first / (float)totalreads * 100.0f;
        fprintf(f, "%s\t%.6f\t%" PRIu64 "\t%" PRIu64, label, percent, e.first, e.second);
        if (o->full_path || specified_ranks) {
            uint64_t id = e.second; std::deque<std::string> lineage; std::map<std::string, std::string> cur;
            if (specified_ranks) for (const auto& r : ranks_list) cur.emplace(r, "NA");
            for (auto it = nodes.find(id); it != nodes.end() && it->second != id; it = nodes.find(id)) {
                if (specified_ranks) {
                    auto rk = node2rank.find(id);
                    if (rk != node2rank.end() && rk->second != "no rank" && ranks_set.count(rk->second)) cur[rk->second] = taxon_name(node2name, id, names_path);
                } else lineage.push_front(taxon_name(node2name, id, names_path));
                id = it->second;
            }
            fprintf(f, "\t");
            if (specified_ranks) for (const auto& r : ranks_list) fprintf(f, "%s;", cur[r].c_str());
            else for (const auto& s : lineage) fprintf(f, "%s;", s.c_str());
        } else fprintf(f, "\t%s", taxon_name(node2name, e.second, names_path).c_str());
        fprintf(f, "\n");
    }
    if (!o->expand_viruses) {
        const float pv = total_virus_reads > 0 ? (float)total_virus_reads / (float)totalreads * 100.0 : 0.0;
        fprintf(f, "%s\t%.6f\t%" PRIu64 "\t%" PRIu64 "\tViruses\n", label, pv, total_virus_reads, kViruses);
    }
    {
        const float pa = above > 0 ? (float)above / (float)totalreads * 100.0 : 0.0;
        fprintf(f, "%s\t%.6f\t%" PRIu64 "\tNA\tcannot be assigned to a (non-viral) %s\n", label, pa, above, rank.c_str());
    }
    if (min_read_count > 0) fprintf(f, "%s\t%.6f\t%" PRIu64 "\tNA\tbelong to a (non-viral) %s having less than %i reads\n", label, (float)below_count / (float)totalreads * 100.0, below_count, rank.c_str(), min_read_count);
    if (min_percent > 0.0) fprintf(f, "%s\t%.6f\t%" PRIu64 "\tNA\tbelong to a (non-viral) %s with less than %g%% of all reads\n", label, (float)below_percent / (float)totalreads * 100.0, below_percent, rank.c_str(), min_percent);
    if (o->filter_unclassified) fprintf(f, "%s\t%.6f\t%" PRIu64 "\tNA\tunclassified\n", label, (float)unclassified / (float)(totalreads + unclassified) * 100.0, unclassified);
    else fprintf(f, "%s\t%.6f\t%" PRIu64 "\tNA\tunclassified\n", label, (float)unclassified / (float)(totalreads) * 100.0, unclassified);
    if (fclose(f) != 0) { kj_err() = std::string("write error on ") + out_path; return KJ_ERR_IO; }
    return KJ_OK;
}
