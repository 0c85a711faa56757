// Minimal safetensors reader: 8-byte LE header length, JSON header {name: {dtype, shape, data_offsets}}, raw data.
// Replaces the reference's SafeTensorFile / safetensor_parse (src/context.h:69-159, src/safetensor.cpp) for the
// tensors the generation path needs. Written from the format description, mmap-free (plain fread).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace ptts_host {

struct StEntry { std::string dtype; std::vector<int64_t> shape; int64_t begin = 0, end = 0; };

class SafeTensors {
public:
    std::map<std::string, StEntry> entries;
    std::string path;
    int64_t data_base = 0;

    bool open(const std::string& p) {
        path = p;
        FILE* f = fopen(p.c_str(), "rb");
        if (!f) return false;
        uint64_t n = 0;
        if (fread(&n, 8, 1, f) != 1 || n == 0 || n > (1ull << 30)) { fclose(f); return false; }
        std::string js(n, '\0');
        if (fread(&js[0], 1, n, f) != n) { fclose(f); return false; }
        data_base = 8 + (int64_t)n;
        int64_t data_size = -1;
        if (fseek(f, 0, SEEK_END) == 0) data_size = (int64_t)ftell(f) - data_base;
        fclose(f);
        pos_ = 0; s_ = &js;
        if (!parse_header() || data_size < 0) return false;
        // Reject malformed / truncated files up front: every consumer sizes its copies from the shape product
        // (the reference trusts the file, src/safetensor.cpp; a user-supplied voice file must not cause an over-read).
        for (auto& kv : entries) {
            const StEntry& e = kv.second;
            const int64_t esz = dtype_size(e.dtype);
            int64_t count = 1;
            for (int64_t d : e.shape) { if (d < 0 || (d > 0 && count > (int64_t)1 << 40)) return false; count *= d; }
            if (e.begin < 0 || e.end < e.begin || e.end > data_size) return false;
            if (esz > 0 && e.end - e.begin != count * esz) return false;
        }
        return true;
    }
    static int64_t dtype_size(const std::string& d) {
        if (d == "F32" || d == "I32" || d == "U32") return 4;
        if (d == "BF16" || d == "F16" || d == "I16" || d == "U16") return 2;
        if (d == "F64" || d == "I64" || d == "U64") return 8;
        if (d == "I8" || d == "U8" || d == "BOOL" || d == "F8_E4M3" || d == "F8_E5M2") return 1;
        return 0;   // unknown dtype: size unchecked, the loader skips it anyway
    }

    // Reads one tensor's raw bytes (file dtype).
    bool read(const StEntry& e, std::vector<uint8_t>& out) const {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        out.resize((size_t)(e.end - e.begin));
        bool ok = fseek(f, data_base + e.begin, SEEK_SET) == 0 && fread(out.data(), 1, out.size(), f) == out.size();
        fclose(f);
        return ok;
    }

private:
    const std::string* s_ = nullptr; size_t pos_ = 0;
    void ws() { while (pos_ < s_->size() && strchr(" \t\r\n", (*s_)[pos_])) pos_++; }
    bool eat(char c) { ws(); if (pos_ < s_->size() && (*s_)[pos_] == c) { pos_++; return true; } return false; }
    bool str(std::string& out) {
        ws(); if (pos_ >= s_->size() || (*s_)[pos_] != '"') return false;
        pos_++; out.clear();
        while (pos_ < s_->size() && (*s_)[pos_] != '"') {
            char c = (*s_)[pos_++];
            if (c == '\\' && pos_ < s_->size()) {
                char d = (*s_)[pos_++];
                switch (d) { case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'u': pos_ += 4; c = '?'; break; default: c = d; }
            }
            out += c;
        }
        if (pos_ >= s_->size()) return false;
        pos_++; return true;
    }
    bool num(int64_t& v) {
        ws(); size_t b = pos_;
        while (pos_ < s_->size() && (isdigit((unsigned char)(*s_)[pos_]) || (*s_)[pos_] == '-')) pos_++;
        if (b == pos_) return false;
        v = strtoll(s_->substr(b, pos_ - b).c_str(), nullptr, 10); return true;
    }
    bool skip_value() {   // skip any JSON value (used for __metadata__ and unknown keys)
        ws(); if (pos_ >= s_->size()) return false;
        char c = (*s_)[pos_];
        if (c == '"') { std::string t; return str(t); }
        if (c == '{' || c == '[') {
            char close = c == '{' ? '}' : ']'; pos_++;
            if (eat(close)) return true;
            do { if (c == '{') { std::string k; if (!str(k) || !eat(':')) return false; } if (!skip_value()) return false; } while (eat(','));
            return eat(close);
        }
        while (pos_ < s_->size() && !strchr(",}] \t\r\n", (*s_)[pos_])) pos_++;
        return true;
    }
    bool num_array(std::vector<int64_t>& v) {
        v.clear(); if (!eat('[')) return false;
        if (eat(']')) return true;
        do { int64_t x; if (!num(x)) return false; v.push_back(x); } while (eat(','));
        return eat(']');
    }
    bool parse_header() {
        if (!eat('{')) return false;
        if (eat('}')) return true;
        do {
            std::string key; if (!str(key) || !eat(':')) return false;
            if (key == "__metadata__") { if (!skip_value()) return false; continue; }
            StEntry e; if (!eat('{')) return false;
            do {
                std::string k; if (!str(k) || !eat(':')) return false;
                if (k == "dtype") { if (!str(e.dtype)) return false; }
                else if (k == "shape") { if (!num_array(e.shape)) return false; }
                else if (k == "data_offsets") { std::vector<int64_t> o; if (!num_array(o) || o.size() != 2) return false; e.begin = o[0]; e.end = o[1]; }
                else if (!skip_value()) return false;
            } while (eat(','));
            if (!eat('}')) return false;
            entries[key] = e;
        } while (eat(','));
        return eat('}');
    }
};

}  // namespace ptts_host
