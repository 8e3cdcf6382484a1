// safetensors.hpp — minimal reader for Hugging Face checkpoints (*.safetensors, single file or a directory of shards).
// File = u64 little-endian header length, JSON header {"name": {"dtype": "BF16", "shape": [..], "data_offsets": [b, e]}, ...},
// raw tensor bytes.  Files are mmap'ed; tensors are handed out as (pointer, dtype, shape) views.
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace oa {

struct StTensor { const uint8_t* data = nullptr; std::string dtype; std::vector<int64_t> shape; size_t bytes = 0; };

class SafeTensors {
public:
    explicit SafeTensors(const std::string& path) {
        struct stat st;
        if (stat(path.c_str(), &st) != 0) throw std::runtime_error("weights path not found: " + path);
        if (S_ISDIR(st.st_mode)) {
            DIR* d = opendir(path.c_str());
            if (!d) throw std::runtime_error("cannot open weights directory: " + path);
            std::vector<std::string> files;
            while (dirent* e = readdir(d)) { std::string n = e->d_name; if (n.size() > 12 && n.substr(n.size() - 12) == ".safetensors") files.push_back(path + "/" + n); }
            closedir(d);
            if (files.empty()) throw std::runtime_error("no *.safetensors files in " + path);
            for (auto& f : files) open_file(f);
        } else open_file(path);
    }
    ~SafeTensors() { for (auto& m : maps_) munmap(m.first, m.second); }
    SafeTensors(const SafeTensors&) = delete;
    bool has(const std::string& name) const { return tensors_.count(name) != 0; }
    const StTensor& get(const std::string& name) const {
        auto it = tensors_.find(name);
        if (it == tensors_.end()) throw std::runtime_error("checkpoint has no tensor '" + name + "'");
        return it->second;
    }

    static size_t dtype_size(const std::string& d) {
        if (d == "BF16" || d == "F16" || d == "I16" || d == "U16") return 2;
        if (d == "F32" || d == "I32" || d == "U32") return 4;
        if (d == "F64" || d == "I64" || d == "U64") return 8;
        if (d == "I8" || d == "U8" || d == "BOOL" || d == "F8_E4M3" || d == "F8_E5M2") return 1;
        return 0;
    }

private:
    void open_file(const std::string& f) {
        int fd = open(f.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + f);
        struct stat st; fstat(fd, &st);
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) throw std::runtime_error("mmap failed for " + f);
        maps_.push_back({m, (size_t)st.st_size});
        const uint8_t* p = reinterpret_cast<const uint8_t*>(m);
        if (st.st_size < 8) throw std::runtime_error("truncated safetensors file " + f);
        uint64_t hl; std::memcpy(&hl, p, 8);
        if (8 + hl > (uint64_t)st.st_size) throw std::runtime_error("bad safetensors header length in " + f);
        parse_header(std::string(reinterpret_cast<const char*>(p + 8), (size_t)hl), p + 8 + hl, (size_t)st.st_size - 8 - (size_t)hl, f);
    }
    // tiny JSON walker for the header's fixed shape
    struct Cur { const std::string& s; size_t i; };
    static void ws(Cur& c) { while (c.i < c.s.size() && (c.s[c.i] == ' ' || c.s[c.i] == '\n' || c.s[c.i] == '\t' || c.s[c.i] == '\r')) ++c.i; }
    static void expect(Cur& c, char ch) { ws(c); if (c.i >= c.s.size() || c.s[c.i] != ch) throw std::runtime_error(std::string("safetensors header: expected '") + ch + "'"); ++c.i; }
    static std::string str(Cur& c) {
        expect(c, '"'); std::string o;
        while (c.i < c.s.size() && c.s[c.i] != '"') { if (c.s[c.i] == '\\' && c.i + 1 < c.s.size()) ++c.i; o += c.s[c.i++]; }
        expect(c, '"'); return o;
    }
    static void skip_value(Cur& c) {
        ws(c);
        if (c.s[c.i] == '"') { str(c); return; }
        if (c.s[c.i] == '{' || c.s[c.i] == '[') {
            const char open = c.s[c.i], close = open == '{' ? '}' : ']'; int depth = 0;
            while (c.i < c.s.size()) { char ch = c.s[c.i]; if (ch == '"') { str(c); continue; } if (ch == open) ++depth; if (ch == close && --depth == 0) { ++c.i; return; } ++c.i; }
            return;
        }
        while (c.i < c.s.size() && c.s[c.i] != ',' && c.s[c.i] != '}' && c.s[c.i] != ']') ++c.i;
    }
    static std::vector<int64_t> int_array(Cur& c) {
        std::vector<int64_t> v; expect(c, '['); ws(c);
        if (c.s[c.i] == ']') { ++c.i; return v; }
        while (true) { ws(c); size_t b = c.i; while (c.i < c.s.size() && (isdigit((unsigned char)c.s[c.i]) || c.s[c.i] == '-')) ++c.i; v.push_back(std::stoll(c.s.substr(b, c.i - b))); ws(c); if (c.s[c.i] == ',') { ++c.i; continue; } expect(c, ']'); break; }
        return v;
    }
    void parse_header(const std::string& h, const uint8_t* data, size_t data_bytes, const std::string& f) {
        Cur c{h, 0}; expect(c, '{'); ws(c);
        if (c.s[c.i] == '}') return;
        while (true) {
            std::string name = str(c); expect(c, ':'); ws(c);
            if (name == "__metadata__") skip_value(c);
            else {
                StTensor t; std::vector<int64_t> off;
                expect(c, '{');
                while (true) {
                    std::string k = str(c); expect(c, ':');
                    if (k == "dtype") t.dtype = str(c); else if (k == "shape") t.shape = int_array(c); else if (k == "data_offsets") off = int_array(c); else skip_value(c);
                    ws(c); if (c.s[c.i] == ',') { ++c.i; continue; } expect(c, '}'); break;
                }
                if (off.size() != 2 || off[0] < 0 || off[1] < off[0] || (size_t)off[1] > data_bytes) throw std::runtime_error("bad data_offsets for " + name + " in " + f);
                t.data = data + off[0]; t.bytes = (size_t)(off[1] - off[0]);
                {   // the byte range must hold exactly prod(shape) elements of the declared dtype: a truncated or mislabelled tensor is an error here, not a SIGBUS later
                    const size_t esz = dtype_size(t.dtype);
                    if (esz == 0) throw std::runtime_error("unsupported dtype " + t.dtype + " for " + name + " in " + f);
                    unsigned __int128 n = 1; for (int64_t d : t.shape) { if (d < 0) throw std::runtime_error("negative dimension for " + name + " in " + f); n *= (unsigned __int128)d; }
                    if (n * esz != (unsigned __int128)t.bytes) throw std::runtime_error("tensor " + name + " in " + f + ": data_offsets span " + std::to_string(t.bytes) + " bytes but shape x dtype needs " + std::to_string((unsigned long long)(n * esz)));
                }
                tensors_[name] = t;
            }
            ws(c); if (c.s[c.i] == ',') { ++c.i; continue; } expect(c, '}'); break;
        }
    }
    std::map<std::string, StTensor> tensors_;
    std::vector<std::pair<void*, size_t>> maps_;
};

}  // namespace oa
