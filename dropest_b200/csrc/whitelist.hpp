// whitelist.hpp -- host side of the whitelist ("real barcodes") merge: file loading, barcode splitting, and the EXACT
// neighbour enumeration used (a) when the device fast path does not apply and (b) to replay order-dependent decisions.
//
// Semantics restated from the reference (our own code; nothing copied):
//   * file format / reverse complement on load ........ BarcodesParser::read_line            BarcodesParser.cpp:117-144
//   * const-length parts, split by cumulative lengths .. ConstLengthBarcodesParser            ConstLengthBarcodesParser.cpp:23-69
//   * inDrop: exactly two lines, split = rest | last len(part2) ... InDropBarcodesParser      InDropBarcodesParser.cpp:15-48
//   * per-part Levenshtein lists sorted by value, depth-first product with sum <= 5 ........ BarcodesParser.cpp:21-74
//   * class walk with fall-through ...................... RealBarcodesMergeStrategy          RealBarcodesMergeStrategy.cpp:63-109
// The reference sorts with std::sort (unstable); where the resulting ORDER can influence a result we call std::sort on
// the same sequence with the same comparison, which reproduces the permutation under the same libstdc++.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dge
{

// Tools::edit_distance (Tools/UtilFunctions.cpp:32-65): single-column Levenshtein with an optional band of half-width
// max_ed around the diagonal, 'N' as a wildcard when skip_n.  Cells outside the band keep stale values and the function
// returns early with the running row minimum (+ off-diagonal penalty) once that exceeds max_ed -- kept literally.
inline unsigned edit_distance_ref(const char *a, const char *b, bool skip_n = true, unsigned max_ed = 10000)
{
    const int la = int(std::strlen(a)), lb = int(std::strlen(b));
    std::vector<int> col(size_t(la) + 1);
    for (int i = 0; i <= la; ++i) col[size_t(i)] = i;
    const int band = int(max_ed);
    for (int j = 1; j <= lb; ++j)
    {
        const int first = std::max(0, j - band), last = std::min(la, j + band);
        int diag = col[size_t(first)];
        col[size_t(first)] = j;
        int row_best = j;
        for (int i = first + 1; i <= last; ++i)
        {
            const int above = col[size_t(i)];
            const char ca = a[i - 1], cb = b[j - 1];
            const bool same = ca == cb || (skip_n && (ca == 'N' || cb == 'N'));
            const int v = std::min(std::min(above + 1, col[size_t(i) - 1] + 1), diag + (same ? 0 : 1));
            row_best = std::min(row_best, v + std::abs(i - j));
            col[size_t(i)] = v;
            diag = above;
        }
        if (unsigned(row_best) > max_ed) return unsigned(row_best);
    }
    return unsigned(col[size_t(la)]);
}

// Tools::hamming_distance (Tools/UtilFunctions.cpp:67-82)
inline unsigned hamming_distance_ref(const std::string &a, const std::string &b, bool skip_n = true)
{
    if (a.size() != b.size()) throw std::runtime_error("Strings should have equal length");
    unsigned d = 0;
    for (size_t i = 0; i < a.size(); ++i)
        if (a[i] != b[i] && !(skip_n && (a[i] == 'N' || b[i] == 'N'))) ++d;
    return d;
}

inline std::string reverse_complement(const std::string &s)
{
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i)
    {
        char c = s[s.size() - 1 - i], o;
        switch (c) { case 'A': o = 'T'; break; case 'T': o = 'A'; break; case 'G': o = 'C'; break; case 'C': o = 'G'; break;
                     case 'N': o = 'N'; break; default: o = '\0'; }
        r[i] = o;
    }
    return r;
}

inline std::string unpack_seq(uint64_t v, unsigned len)
{
    std::string s(len, 'A');
    for (unsigned i = 0; i < len; ++i) s[len - 1 - i] = "ACGT"[(v >> (2 * i)) & 3];
    return s;
}

inline bool pack_seq(const std::string &s, uint64_t &out)
{
    out = 0;
    for (char c : s)
    {
        unsigned b;
        switch (c) { case 'A': b = 0; break; case 'C': b = 1; break; case 'G': b = 2; break; case 'T': b = 3; break; default: return false; }
        out = (out << 2) | b;
    }
    return true;
}

class Whitelist
{
public:
    static const unsigned MAX_SUM_DISTANCE = 5; // BarcodesParser::MAX_REAL_MERGE_EDIT_DISTANCE (BarcodesParser.h:57)

    std::vector<std::vector<std::string>> parts;
    bool indrop = false;
    size_t indrop_len2 = 0;
    std::vector<size_t> const_lengths;
    size_t const_total = 0;

    bool empty() const { return parts.empty(); }

    void load(const std::string &fname, bool indrop_type)
    {
        indrop = indrop_type;
        std::ifstream f(fname);
        if (f.fail()) throw std::runtime_error("Can't open file with barcodes: '" + fname + "'");
        std::string line;
        if (indrop)
        {
            parts.assign(2, {});
            for (int i = 0; i < 2; ++i)
                if (!std::getline(f, line) || !parse_line(line, parts[size_t(i)], false) || parts[size_t(i)].empty())
                    throw std::runtime_error("File with barcodes (" + fname + ") has wrong format");
        }
        else
        {
            while (std::getline(f, line))
            {
                std::vector<std::string> p;
                parse_line(line, p, true);
                if (p.empty()) throw std::runtime_error("File with barcodes (" + fname + ") has wrong format");
                parts.push_back(p);
            }
        }
        if (parts.empty()) throw std::runtime_error("ERROR: empty barcodes list");
        for (auto const &p : parts)
            if (p.empty()) throw std::runtime_error("ERROR: empty barcodes list");
        if (indrop) indrop_len2 = parts[1][0].size();
        else
            for (auto const &p : parts) { const_lengths.push_back(p[0].size()); const_total += p[0].size(); }
    }

    std::vector<std::string> split(const std::string &cb) const
    {
        std::vector<std::string> res;
        if (indrop)
        {
            res.push_back(cb.substr(0, cb.size() - indrop_len2));
            res.push_back(cb.substr(cb.size() - indrop_len2));
            return res;
        }
        if (cb.size() != const_total)
            throw std::runtime_error("Barcode '" + cb + "' has wrong length (" + std::to_string(const_total) + " expected)");
        size_t pos = 0;
        for (size_t len : const_lengths) { res.push_back(cb.substr(pos, len)); pos += len; }
        return res;
    }

    // Is `cb` itself a whitelist barcode (every part equals a token)?  Such a cell is its own merge target.
    bool contains(const std::string &cb) const
    {
        std::vector<std::string> p;
        try { p = split(cb); } catch (std::exception &) { return false; }
        for (size_t k = 0; k < parts.size(); ++k)
            if (std::find(parts[k].begin(), parts[k].end(), p[k]) == parts[k].end()) return false;
        return true;
    }

    // True when every part has one token length, tokens are N-free, and the lengths add up to cb_len:
    // then distance classes 0/1 are Hamming classes and the device fast path is exact.
    bool fast_path_ok(size_t cb_len) const
    {
        if (parts.empty() || parts.size() > 4) return false;
        size_t total = 0;
        for (auto const &p : parts)
        {
            const size_t len = p[0].size();
            if (len == 0 || len > 16) return false;
            for (auto const &t : p)
            {
                uint64_t tmp;
                if (t.size() != len || !pack_seq(t, tmp)) return false;
            }
            total += len;
        }
        if (indrop && parts.size() != 2) return false;
        return total == cb_len;
    }

    struct Leaf { size_t inds[4]; unsigned sum; };

    // All whitelist combinations with summed per-part distance <= 5, in the reference's traversal order, then ordered by
    // std::sort on the sum exactly as RealBarcodesMergeStrategy.cpp:74-76 does.
    std::vector<Leaf> sorted_leaves(const std::string &cb) const
    {
        struct IV { size_t index; long value; };
        const std::vector<std::string> cb_parts = split(cb);
        std::vector<std::vector<IV>> dists(parts.size());
        for (size_t p = 0; p < parts.size(); ++p)
        {
            for (size_t t = 0; t < parts[p].size(); ++t)
                dists[p].push_back(IV{t, long(edit_distance_ref(cb_parts[p].c_str(), parts[p][t].c_str()))});
            std::sort(dists[p].begin(), dists[p].end(), [](const IV &x, const IV &y) { return x.value < y.value; });
        }
        std::vector<Leaf> leaves;
        Leaf cur{};
        std::function<void(size_t, unsigned)> rec = [&](size_t p, unsigned acc) {
            if (p == parts.size()) { cur.sum = acc; leaves.push_back(cur); return; }
            for (auto const &d : dists[p])
            {
                unsigned s = acc + unsigned(d.value);
                if (s > MAX_SUM_DISTANCE) return;
                cur.inds[p] = d.index;
                rec(p + 1, s);
            }
        };
        rec(0, 0);
        std::sort(leaves.begin(), leaves.end(), [](const Leaf &x, const Leaf &y) { return x.sum < y.sum; });
        return leaves;
    }

    std::string barcode_of(const Leaf &l) const
    {
        std::string s;
        for (size_t p = 0; p < parts.size(); ++p) s += parts[p].at(l.inds[p]);
        return s;
    }

    // Exact neighbour list.  `lookup(barcode) -> id or -1` finds a cell by barcode string; `eligible(id)` applies the
    // size / umis conditions (RealBarcodesMergeStrategy.cpp:98-103).  poisson selects get_max_merge_dist of the Poisson
    // strategy (PoissonRealBarcodesMergeStrategy.cpp:20-23).
    template <class Lookup, class Eligible>
    std::vector<long> neighbours(const std::string &cb, bool poisson, Lookup lookup, Eligible eligible) const
    {
        std::vector<long> res;
        std::vector<Leaf> leaves = sorted_leaves(cb);
        if (leaves.empty()) return res;
        const unsigned min_d = leaves.front().sum;
        unsigned max_dist = poisson ? (min_d == 0 ? 2u : min_d + 1) : min_d;
        for (auto const &leaf : leaves)
        {
            if (leaf.sum > max_dist && !res.empty()) break;
            long id = lookup(barcode_of(leaf));
            if (id >= 0 && eligible(id)) res.push_back(id);
            max_dist = std::max(max_dist, leaf.sum);
        }
        return res;
    }

private:
    static bool parse_line(const std::string &line, std::vector<std::string> &out, bool require_equal_length)
    {
        std::istringstream in(line);
        std::string tok;
        size_t len = 0;
        while (in >> tok)
        {
            if (len == 0) len = tok.size();
            else if (require_equal_length && len != tok.size())
                throw std::runtime_error("All barcodes in one line must have the same length");
            out.push_back(reverse_complement(tok));
        }
        return true;
    }
};

} // namespace dge
