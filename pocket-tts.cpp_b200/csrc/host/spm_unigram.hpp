// SentencePiece *unigram* encoder, dependency-free: parses tokenizer.model (protobuf wire format), applies the
// normaliser (precompiled charsmap = Darts double-array trie + replacement strings, whitespace rules, dummy
// prefix) and runs the Viterbi search with SentencePiece's scoring rules, incl. user-defined symbols and byte
// fallback. Replaces sentencepiece::SentencePieceProcessor::Load/Encode as used by the reference
// (src/pocket_tts/conditioners/text.h:10-27, src/pocket_tts.cpp:8). The algorithm restates SentencePiece
// (google/sentencepiece, un-pinned by the reference) normalizer.cc / unigram_model.cc (EncodeOptimized) /
// sentencepiece_processor.cc from its published behaviour; ids are verified bit-exact against the upstream wheel
// in tests/test_text_frontend.py.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace ptts_host {

class SpmUnigram {
public:
    enum PieceType { NORMAL = 1, UNKNOWN = 2, CONTROL = 3, USER_DEFINED = 4, UNUSED = 5, BYTE = 6 };
    struct Piece { std::string s; float score = 0.f; int type = NORMAL; };

    bool load(const std::string& path) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        std::string blob; char buf[65536]; size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) blob.append(buf, n);
        fclose(f);
        if (!parse_model(blob)) return false;
        build();
        return true;
    }
    int vocab_size() const { return (int)pieces_.size(); }
    const Piece& piece(int id) const { return pieces_[id]; }

    std::vector<int> encode(const std::string& text) const {
        const std::string norm = normalize(text);
        std::vector<int> ids;
        if (norm.empty()) return ids;
        std::vector<std::pair<std::pair<int, int>, int>> res = viterbi(norm);   // ((begin,end), id)
        bool prev_unk = false;
        for (auto& r : res) {
            const int id = r.second; const bool is_unk = id == unk_id_;
            if (is_unk && byte_fallback_) {
                for (int i = r.first.first; i < r.first.second; i++) ids.push_back(byte_id_[(uint8_t)norm[i]]);
            } else if (!(prev_unk && is_unk)) {
                ids.push_back(id);
            }
            prev_unk = is_unk;
        }
        return ids;
    }

    std::string normalize(const std::string& input_s) const {
        const char* in = input_s.data(); size_t len = input_s.size();
        std::string out;
        const std::string kSpace = "\xe2\x96\x81";
        std::string tmp; int consumed = 0;
        if (remove_extra_ws_) {
            while (len > 0) { norm_prefix(in, len, tmp, consumed); if (tmp != " ") break; in += consumed; len -= consumed; }
        }
        if (len == 0) return out;
        auto add_ws = [&]() { if (escape_ws_) out += kSpace; else out += ' '; };
        if (!ws_suffix_ && add_dummy_prefix_) add_ws();
        bool is_prev_space = remove_extra_ws_;
        while (len > 0) {
            norm_prefix(in, len, tmp, consumed);
            size_t b = 0;
            while (is_prev_space && b < tmp.size() && tmp[b] == ' ') b++;
            if (b < tmp.size()) {
                for (size_t i = b; i < tmp.size(); i++) { if (escape_ws_ && tmp[i] == ' ') out += kSpace; else out += tmp[i]; }
                is_prev_space = tmp.back() == ' ';
            }
            in += consumed; len -= consumed;
            if (!remove_extra_ws_) is_prev_space = false;
        }
        if (remove_extra_ws_) {
            const std::string sp = escape_ws_ ? kSpace : std::string(" ");
            while (out.size() >= sp.size() && out.compare(out.size() - sp.size(), sp.size(), sp) == 0) out.resize(out.size() - sp.size());
        }
        if (ws_suffix_ && add_dummy_prefix_) add_ws();
        return out;
    }

private:
    std::vector<Piece> pieces_;
    int unk_id_ = 0; bool byte_fallback_ = false, ws_suffix_ = false;
    bool add_dummy_prefix_ = true, remove_extra_ws_ = true, escape_ws_ = true;
    std::string charsmap_;
    const uint32_t* darts_ = nullptr; size_t darts_n_ = 0; const char* norm_strs_ = nullptr; size_t norm_strs_n_ = 0;
    float min_score_ = FLT_MAX, max_score_ = FLT_MIN;
    int byte_id_[256];
    std::vector<std::string> user_defined_;
    // piece trie: nodes hold sorted (byte, child) edges; value = piece id or -1
    struct Node { std::vector<std::pair<uint8_t, int>> edges; int value = -1; };
    std::vector<Node> trie_;

    // ---------------- protobuf wire format ----------------
    static bool varint(const uint8_t*& p, const uint8_t* e, uint64_t& v) {
        v = 0; int shift = 0;
        while (p < e) { uint8_t b = *p++; v |= (uint64_t)(b & 0x7f) << shift; if (!(b & 0x80)) return true; shift += 7; if (shift > 63) return false; }
        return false;
    }
    template <typename F> static bool fields(const uint8_t* p, const uint8_t* e, F&& cb) {
        while (p < e) {
            uint64_t key; if (!varint(p, e, key)) return false;
            const int field = (int)(key >> 3), wt = (int)(key & 7);
            uint64_t v = 0; const uint8_t* data = nullptr; size_t dl = 0;
            if (wt == 0) { if (!varint(p, e, v)) return false; }
            else if (wt == 1) { if (e - p < 8) return false; memcpy(&v, p, 8); p += 8; }
            else if (wt == 5) { if (e - p < 4) return false; uint32_t w; memcpy(&w, p, 4); v = w; p += 4; }
            else if (wt == 2) { uint64_t l; if (!varint(p, e, l) || (uint64_t)(e - p) < l) return false; data = p; dl = (size_t)l; p += l; }
            else return false;
            cb(field, wt, v, data, dl);
        }
        return true;
    }
    bool parse_model(const std::string& blob) {
        const uint8_t* b = (const uint8_t*)blob.data(); const uint8_t* e = b + blob.size();
        return fields(b, e, [&](int f, int wt, uint64_t, const uint8_t* d, size_t dl) {
            if (wt != 2) return;
            if (f == 1) {                                   // ModelProto.pieces
                Piece pc;
                fields(d, d + dl, [&](int f2, int wt2, uint64_t v2, const uint8_t* d2, size_t dl2) {
                    if (f2 == 1 && wt2 == 2) pc.s.assign((const char*)d2, dl2);
                    else if (f2 == 2 && wt2 == 5) { uint32_t w = (uint32_t)v2; memcpy(&pc.score, &w, 4); }
                    else if (f2 == 3 && wt2 == 0) pc.type = (int)v2;
                });
                pieces_.push_back(pc);
            } else if (f == 2) {                            // trainer_spec
                fields(d, d + dl, [&](int f2, int wt2, uint64_t v2, const uint8_t*, size_t) {
                    if (wt2 != 0) return;
                    if (f2 == 35) byte_fallback_ = v2 != 0;
                    else if (f2 == 40) unk_id_ = (int)(int64_t)v2;
                    else if (f2 == 24) ws_suffix_ = v2 != 0;
                });
            } else if (f == 3) {                            // normalizer_spec
                fields(d, d + dl, [&](int f2, int wt2, uint64_t v2, const uint8_t* d2, size_t dl2) {
                    if (f2 == 2 && wt2 == 2) charsmap_.assign((const char*)d2, dl2);
                    else if (f2 == 3 && wt2 == 0) add_dummy_prefix_ = v2 != 0;
                    else if (f2 == 4 && wt2 == 0) remove_extra_ws_ = v2 != 0;
                    else if (f2 == 5 && wt2 == 0) escape_ws_ = v2 != 0;
                });
            }
        }) && !pieces_.empty();
    }

    void build() {
        for (int i = 0; i < 256; i++) byte_id_[i] = unk_id_;
        trie_.assign(1, Node());
        for (int id = 0; id < (int)pieces_.size(); id++) {
            const Piece& p = pieces_[id];
            if (p.type == NORMAL) { min_score_ = std::min(min_score_, p.score); max_score_ = std::max(max_score_, p.score); }
            if (p.type == BYTE && p.s.size() == 6) byte_id_[strtol(p.s.substr(3, 2).c_str(), nullptr, 16)] = id;
            if (p.type == USER_DEFINED) user_defined_.push_back(p.s);
            if (p.type == NORMAL || p.type == USER_DEFINED || p.type == UNUSED) {
                int node = 0;
                for (unsigned char c : p.s) {
                    int next = -1;
                    for (auto& ed : trie_[node].edges) if (ed.first == c) { next = ed.second; break; }
                    if (next < 0) { next = (int)trie_.size(); trie_[node].edges.push_back({c, next}); trie_.push_back(Node()); }
                    node = next;
                }
                if (trie_[node].value < 0) trie_[node].value = id;
            }
        }
        if (charsmap_.size() >= 4) {
            uint32_t tsz; memcpy(&tsz, charsmap_.data(), 4);
            if (tsz + 4 <= charsmap_.size()) {
                darts_ = (const uint32_t*)(charsmap_.data() + 4); darts_n_ = tsz / 4;
                norm_strs_ = charsmap_.data() + 4 + tsz; norm_strs_n_ = charsmap_.size() - 4 - tsz;
            }
        }
    }

    // ---------------- normaliser ----------------
    static bool trail(char c) { return (signed char)c < -0x40; }
    static bool valid_cp(uint32_t c) { return c < 0xD800 || (c >= 0xE000 && c <= 0x10FFFF); }
    static uint32_t decode_utf8(const char* b, size_t len, size_t* mblen) {
        const unsigned char c0 = (unsigned char)b[0];
        if (c0 < 0x80) { *mblen = 1; return c0; }
        if (len >= 2 && (c0 & 0xE0) == 0xC0) {
            const uint32_t cp = ((c0 & 0x1F) << 6) | (b[1] & 0x3F);
            if (trail(b[1]) && cp >= 0x80 && valid_cp(cp)) { *mblen = 2; return cp; }
        } else if (len >= 3 && (c0 & 0xF0) == 0xE0) {
            const uint32_t cp = ((c0 & 0x0F) << 12) | ((b[1] & 0x3F) << 6) | (b[2] & 0x3F);
            if (trail(b[1]) && trail(b[2]) && cp >= 0x800 && valid_cp(cp)) { *mblen = 3; return cp; }
        } else if (len >= 4 && (c0 & 0xF8) == 0xF0) {
            const uint32_t cp = ((c0 & 0x07) << 18) | ((b[1] & 0x3F) << 12) | ((b[2] & 0x3F) << 6) | (b[3] & 0x3F);
            if (trail(b[1]) && trail(b[2]) && trail(b[3]) && cp >= 0x10000 && valid_cp(cp)) { *mblen = 4; return cp; }
        }
        *mblen = 1; return 0xFFFD;
    }
    // longest charsmap rule at `in` (Darts commonPrefixSearch), else one UTF-8 char; user-defined symbols pass through.
    void norm_prefix(const char* in, size_t len, std::string& out, int& consumed) const {
        size_t best_uds = 0;
        for (auto& u : user_defined_) if (u.size() > best_uds && u.size() <= len && memcmp(in, u.data(), u.size()) == 0) best_uds = u.size();
        if (best_uds) { out.assign(in, best_uds); consumed = (int)best_uds; return; }
        size_t longest = 0; uint32_t value = 0;
        if (darts_) {
            auto offset = [](uint32_t u) { return (u >> 10) << ((u & (1u << 9)) >> 6); };
            size_t node = 0; uint32_t unit = darts_[0]; node ^= offset(unit);
            for (size_t i = 0; i < len; i++) {
                const uint8_t c = (uint8_t)in[i];
                node ^= c;
                if (node >= darts_n_) break;
                unit = darts_[node];
                if ((unit & ((1u << 31) | 0xFF)) != c) break;
                node ^= offset(unit);
                if ((unit >> 8) & 1) {
                    if (node >= darts_n_) break;
                    const uint32_t v = darts_[node] & ((1u << 31) - 1);
                    if (longest == 0 || i + 1 > longest) { longest = i + 1; value = v; }
                }
            }
        }
        if (longest == 0) {
            size_t mblen = 0; const uint32_t cp = decode_utf8(in, len, &mblen);
            if (cp == 0xFFFD && mblen != 3) { consumed = 1; out = "\xEF\xBF\xBD"; }
            else { consumed = (int)mblen; out.assign(in, mblen); }
        } else {
            consumed = (int)longest;
            out = (value < norm_strs_n_) ? std::string(norm_strs_ + value) : std::string();
        }
    }

    // ---------------- Viterbi (unigram_model.cc EncodeOptimized) ----------------
    static int one_char_len(char c) { return "\1\1\1\1\1\1\1\1\1\1\1\1\2\2\3\4"[((unsigned char)c) >> 4]; }
    std::vector<std::pair<std::pair<int, int>, int>> viterbi(const std::string& norm) const {
        struct Best { int id = -1; float score = 0.f; int starts_at = -1; };
        const int size = (int)norm.size();
        const float unk_score = min_score_ - 10.0f;
        std::vector<Best> best(size + 1);
        int starts_at = 0;
        while (starts_at < size) {
            const float till_here = best[starts_at].score;
            bool has_single = false;
            const int mblen = std::min(one_char_len(norm[starts_at]), size - starts_at);
            int node = 0;
            for (int key_pos = starts_at; key_pos < size;) {
                const unsigned char c = (unsigned char)norm[key_pos];
                int next = -1;
                for (auto& ed : trie_[node].edges) if (ed.first == c) { next = ed.second; break; }
                if (next < 0) break;
                node = next; key_pos++;
                const int id = trie_[node].value;
                if (id >= 0) {
                    if (pieces_[id].type == UNUSED) continue;
                    Best& t = best[key_pos];
                    const size_t length = (size_t)(key_pos - starts_at);
                    const double score = pieces_[id].type == USER_DEFINED ? ((float)length * max_score_ - 0.1) : (double)pieces_[id].score;
                    const double cand = score + till_here;
                    if (t.starts_at == -1 || cand > t.score) { t.score = (float)cand; t.starts_at = starts_at; t.id = id; }
                    if (!has_single && (int)length == mblen) has_single = true;
                }
            }
            if (!has_single) {
                Best& t = best[starts_at + mblen];
                const float cand = unk_score + till_here;
                if (t.starts_at == -1 || cand > t.score) { t.score = cand; t.starts_at = starts_at; t.id = unk_id_; }
            }
            starts_at += mblen;
        }
        std::vector<std::pair<std::pair<int, int>, int>> res;
        int ends_at = size;
        while (ends_at > 0) { const Best& n = best[ends_at]; res.push_back({{n.starts_at, ends_at}, n.id}); ends_at = n.starts_at; }
        std::reverse(res.begin(), res.end());
        return res;
    }
};

}  // namespace ptts_host
