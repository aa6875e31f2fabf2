// ASAN/UBSAN driver: reads a corpus file of length-prefixed records (kind byte, u32 length, payload) and feeds the readers.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "ptx_emit.h"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    std::vector<uint8_t> buf; uint8_t tmp[65536]; size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    size_t i = 0, ok = 0, err = 0;
    while (i + 5 <= buf.size()) {
        uint8_t kind = buf[i]; uint32_t len; memcpy(&len, &buf[i + 1], 4); i += 5;
        if (i + len > buf.size()) break;
        try {
            if (kind == 0) { std::vector<uint32_t> w(len / 4); if (len >= 4) memcpy(w.data(), &buf[i], len / 4 * 4); kir::disassemble_spirv(w.data(), w.size()); }
            else if (kind == 1) { std::vector<uint32_t> w(len / 4); if (len >= 4) memcpy(w.data(), &buf[i], len / 4 * 4); uint32_t ls[3]; kir::translate_spirv(w.data(), w.size(), "main", 5, ls); }
            else if (kind == 4) { std::vector<uint32_t> w(len / 4); if (len >= 4) memcpy(w.data(), &buf[i], len / 4 * 4); uint32_t ls[3]; auto m = kir::translate_spirv_v2(w.data(), w.size(), "main", 0, ls); kir::module_to_text(*m); kir::emit_ptx(*m, nullptr); }
            else if (kind == 5) { auto m = kir::parse_module(std::string((const char*)&buf[i], len), 2); kir::module_to_text(*m); kir::emit_ptx(*m, nullptr); }
            else if (kind == 2) { auto m = kir::parse_module(std::string((const char*)&buf[i], len)); kir::module_to_text(*m); kir::emit_ptx(*m, nullptr); }
            else { size_t z = 0; while (z < len && buf[i + z]) ++z; std::string k((const char*)&buf[i], z); if (z < len) kir::round_trip(k, std::string((const char*)&buf[i + z + 1], len - z - 1)); }
            ++ok;
        } catch (const kir::Error&) { ++err; } catch (const std::exception&) { ++err; }
        i += len;
    }
    printf("ok %zu err %zu\n", ok, err);
}
