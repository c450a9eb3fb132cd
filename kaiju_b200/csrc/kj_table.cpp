// kj_table.cpp -- kaiju2table's summary report (the observable contract of src/kaiju2table.cpp:150-365) written from per-taxon read counts
// (the vector a context accumulates in HBM, kj_counts_get) instead of re-reading the per-read output file.
//
// Own design: the taxonomy is loaded into DENSE arrays (ids sorted, parent / rank / name by index), nodes are ordered by depth, and the
// report follows from two sweeps over that order -- top-down: "lies below Viruses (taxon 10239)"; bottom-up: reads of a taxon plus the
// reads of its non-viral descendants (the reference's per-hit ancestor walk, summed the other way round).  What has to match the
// reference byte for byte is the output: the row set, the order (descending reads, ties by ascending id), the float expressions of the
// percentages and the printf formats; tests/test_table.py pins that against the unmodified tool.
// Observable rules kept: nodes.dmp / names.dmp are parsed with the reference's tolerant rules (util.cpp:123-178); reads of taxa missing
// from nodes.dmp only count towards the total (with a warning); viral taxa are listed on their own, never rolled up; the root never gets
// a row; -m / -c thresholds, -u, -e, -p / -l as in the tool.
#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <set>
#include <vector>
#include "kj_host.h"

namespace {
const uint64_t kViruses = 10239;

bool digits(const std::string& s, size_t b, size_t e, uint64_t& v) {
    if (b == std::string::npos || b >= s.size()) return false;
    if (e == std::string::npos) e = s.size();
    if (e <= b) return false;
    v = 0; for (size_t i = b; i < e; i++) { if (s[i] < '0' || s[i] > '9') return false; v = v * 10 + (uint64_t)(s[i] - '0'); }
    return true;
}
struct Taxonomy {
    std::vector<uint64_t> id;            // ascending
    std::vector<int32_t> parent;         // index of the parent; own index for a self-parent root; -1 if the parent is not in nodes.dmp
    std::vector<std::string> rank, name; std::vector<uint8_t> has_name;
    std::vector<uint32_t> order;         // indices by increasing depth (parents before children)
    std::vector<uint8_t> viral;          // at or below Viruses
    int find(uint64_t x) const { auto it = std::lower_bound(id.begin(), id.end(), x); return (it != id.end() && *it == x) ? (int)(it - id.begin()) : -1; }
};
int load_taxonomy(const char* nodes_path, const char* names_path, Taxonomy& T) {
    struct Row { uint64_t id, parent; std::string rank; };
    std::vector<Row> rows;
    {
        std::ifstream f(nodes_path); if (!f.is_open()) { kj_err() = std::string("Could not open file ") + nodes_path; return KJ_ERR_IO; }
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty()) continue;
            size_t end = line.find_first_not_of("0123456789"); uint64_t node, par;
            if (!digits(line, 0, end, node)) continue;
            size_t start = line.find_first_of("0123456789", end); if (start == std::string::npos) continue;
            end = line.find_first_not_of("0123456789", start + 1);
            if (!digits(line, start, end, par)) continue;
            start = line.find_first_of("abcdefghijklmnopqrstuvwxyz", end);
            if (start == std::string::npos) continue;              // a line without a rank is dropped by the reference as well
            size_t e2 = line.find_first_not_of("abcdefghijklmnopqrstuvwxyz ", start);
            rows.push_back({node, par, line.substr(start, e2 == std::string::npos ? std::string::npos : e2 - start)});
        }
    }
    // first occurrence of an id wins (the reference's emplace): stable sort, then unique
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.id < b.id; });
    rows.erase(std::unique(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.id == b.id; }), rows.end());
    const size_t n = rows.size();
    T.id.resize(n); T.parent.resize(n); T.rank.resize(n); T.name.assign(n, std::string()); T.has_name.assign(n, 0);
    for (size_t i = 0; i < n; i++) { T.id[i] = rows[i].id; T.rank[i].swap(rows[i].rank); }
    for (size_t i = 0; i < n; i++) T.parent[i] = T.find(rows[i].parent);
    {
        std::ifstream f(names_path); if (!f.is_open()) { kj_err() = std::string("Could not open file ") + names_path; return KJ_ERR_IO; }
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty() || line.find("scientific name") == std::string::npos) continue;
            size_t start = line.find_first_of("0123456789"); if (start == std::string::npos) continue;
            size_t end = line.find_first_not_of("0123456789", start); uint64_t x;
            if (!digits(line, start, end, x)) continue;
            start = line.find_first_not_of("\t|", end); if (start == std::string::npos) continue;
            end = line.find_first_of("\t|", start + 1);
            const int i = T.find(x);
            if (i >= 0 && !T.has_name[(size_t)i]) { T.name[(size_t)i] = line.substr(start, end == std::string::npos ? std::string::npos : end - start); T.has_name[(size_t)i] = 1; }
        }
    }
    // depth by memoised climbing (a missing parent or a cycle ends the climb); order = counting sort by depth
    std::vector<int32_t> depth(n, -1); std::vector<uint8_t> visiting(n, 0); std::vector<uint32_t> path;
    for (size_t i = 0; i < n; i++) {
        if (depth[i] >= 0) continue;
        path.clear(); size_t cur = i; int32_t d = 0;
        for (;;) {
            if (depth[cur] >= 0) { d = depth[cur]; break; }
            if (visiting[cur]) { d = 0; break; }
            const int32_t p = T.parent[cur];
            if (p < 0 || (size_t)p == cur) { depth[cur] = 0; d = 0; break; }
            visiting[cur] = 1; path.push_back((uint32_t)cur); cur = (size_t)p;
        }
        for (size_t k = path.size(); k-- > 0;) depth[path[k]] = ++d;
    }
    int32_t maxd = 0; for (size_t i = 0; i < n; i++) maxd = std::max(maxd, depth[i]);
    std::vector<uint32_t> start((size_t)maxd + 2, 0);
    for (size_t i = 0; i < n; i++) start[(size_t)depth[i] + 1]++;
    for (size_t d = 1; d < start.size(); d++) start[d] += start[d - 1];
    T.order.resize(n); for (size_t i = 0; i < n; i++) T.order[start[(size_t)depth[i]]++] = (uint32_t)i;
    T.viral.assign(n, 0);
    for (uint32_t i : T.order) { const int32_t p = T.parent[i]; T.viral[i] = (T.id[i] == kViruses || (p >= 0 && p != (int32_t)i && T.viral[(size_t)p])) ? 1 : 0; }
    return KJ_OK;
}
std::string display_name(const Taxonomy& T, uint32_t i, const char* names_path) {
    if (T.has_name[i]) return T.name[i];
    fprintf(stderr, "Warning: Taxon ID %" PRIu64 " is not found in file %s.\n", T.id[i], names_path);
    return "taxonid:" + std::to_string(T.id[i]);
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
    std::vector<std::string> ranks_list; std::set<std::string> ranks_set;
    if (specified_ranks) {
        const std::string a = o->rank_list; size_t b = 0;
        while (b <= a.size()) { size_t e = a.find(',', b); if (e == std::string::npos) e = a.size(); if (e > b) { ranks_list.push_back(a.substr(b, e - b)); ranks_set.insert(a.substr(b, e - b)); } b = e + 1; }
        if (!ranks_set.count(rank)) { kj_err() = "Specified rank " + rank + " is not contained in rank list supplied with option -l"; return KJ_ERR_ARG; }
    }
    Taxonomy T;
    int rc = load_taxonomy(nodes_path, names_path, T); if (rc) return rc;
    const size_t nt = T.id.size();
    if (T.find(kViruses) < 0) fprintf(stderr, "Taxon ID %" PRIu64 " not found in taxonomy!\n", kViruses);

    // reads per taxon index; reads of unknown taxa and unclassified reads only enter the totals
    std::vector<uint64_t> direct(nt, 0); uint64_t unclassified = 0, totalreads = 0, total_virus_reads = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t c = counts[i]; if (!c) continue;
        totalreads += c;
        if (taxon_ids[i] == 0) { unclassified += c; continue; }
        const int k = T.find(taxon_ids[i]);
        if (k < 0) { fprintf(stderr, "Warning: Taxon ID %" PRIu64 " is not contained in %s.\n", taxon_ids[i], nodes_path); continue; }
        direct[(size_t)k] += c;
        if (T.viral[(size_t)k]) total_virus_reads += c;
    }
    // bottom-up: a non-viral taxon holds its own reads and those of its descendants; viral taxa keep their own reads only
    std::vector<uint64_t> reads(direct);
    for (size_t k = nt; k-- > 0;) {
        const uint32_t i = T.order[k]; const int32_t p = T.parent[i];
        if (!T.viral[i] && reads[i] && p >= 0 && p != (int32_t)i) reads[(size_t)p] += reads[i];
    }
    if (o->filter_unclassified) totalreads -= unclassified;
    // rows: viral taxa with reads (always), taxa of the requested rank above the thresholds; a self-parent root never gets a row
    struct RowOut { uint64_t count; uint32_t idx; };
    std::vector<RowOut> out_rows; uint64_t at_rank = 0, below_percent = 0, below_count = 0;
    for (uint32_t i = 0; i < nt; i++) {                       // ascending id: equal counts keep this order (stable sort below)
        if (T.viral[i]) { if (direct[i]) out_rows.push_back({direct[i], i}); continue; }
        if (!reads[i] || T.parent[i] == (int32_t)i || T.rank[i] != rank) continue;
        const uint64_t count = reads[i];
        if ((int)count >= min_read_count) {
            const float percent = (float)count / (float)totalreads * 100;
            if (percent >= min_percent) out_rows.push_back({count, i}); else below_percent += count;
        } else below_count += count;
        at_rank += count;
    }
    std::stable_sort(out_rows.begin(), out_rows.end(), [](const RowOut& a, const RowOut& b) { return a.count > b.count; });
    uint64_t above = o->filter_unclassified ? totalreads - at_rank : totalreads - unclassified - at_rank;
    above -= total_virus_reads;

    FILE* f = fopen(out_path, append ? "a" : "w");
    if (!f) { kj_err() = std::string("Could not open file ") + out_path + " for writing"; return KJ_ERR_IO; }
    if (!append) fprintf(f, "file\tpercent\treads\ttaxon_id\ttaxon_name\n");
    std::vector<std::string> ranks_vec(ranks_list.begin(), ranks_list.end()), field;
    for (const RowOut& e : out_rows) {
        if (!o->expand_viruses && T.viral[e.idx]) continue;
        const float percent = (float)e.count / (float)totalreads * 100.0f;
        fprintf(f, "%s\t%.6f\t%" PRIu64 "\t%" PRIu64, label, percent, e.count, T.id[e.idx]);
        if (o->full_path || specified_ranks) {
            // the lineage from the taxon up to (not including) the root / a missing parent
            std::vector<uint32_t> lin; for (int32_t i = (int32_t)e.idx; i >= 0 && T.parent[(size_t)i] != i && lin.size() <= nt; i = T.parent[(size_t)i]) lin.push_back((uint32_t)i);
            fprintf(f, "\t");
            if (specified_ranks) {
                field.assign(ranks_vec.size(), "NA");
                for (uint32_t i : lin) {                              // a rank occurring twice in a lineage: the node nearer to the root is reported
                    if (T.rank[i] == "no rank" || !ranks_set.count(T.rank[i])) continue;
                    const std::string nm = display_name(T, i, names_path);
                    for (size_t r = 0; r < ranks_vec.size(); r++) if (ranks_vec[r] == T.rank[i]) field[r] = nm;
                }
                for (const auto& x : field) fprintf(f, "%s;", x.c_str());
            } else for (size_t k = lin.size(); k-- > 0;) fprintf(f, "%s;", display_name(T, lin[k], names_path).c_str());
        } else fprintf(f, "\t%s", display_name(T, e.idx, names_path).c_str());
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
