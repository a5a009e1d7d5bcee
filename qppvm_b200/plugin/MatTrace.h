// MatTrace.h -- MatLogger-compatible trace of the per-tick solver outputs.
//
// The reference dumps its per-tick quantities through XBot::MatLogger ("tau_qp", "tau_desired", "time_matlogger":
// ref:src/QPPVMPlugin.cpp:250-258,322-325; "<link>_wrench", "tau", "tau_c", "qddot_value", "x":
// ref:src/ForceAcc.cpp:200,233-236; flushed in close(): ref:src/QPPVMPlugin.cpp:341, ref:include/ForceAccPlugin/ForceAcc.h:43)
// and inspects them offline in MATLAB.  This header-only writer keeps the same contract -- add(name, vector) once per
// tick, flush() at close -- and writes a MATLAB Level-5 MAT-file in which every name is an n x T double matrix (one
// column per tick), which is what MatLogger produces.  Used by the drop-in plugins (through the XBot::MatLogger of the
// runtime, or the test shim's) and usable on its own for batched runs.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace qppvm {

class MatTrace {
public:
    void add(const std::string& name, const double* v, int n)
    {
        Series& s = series_[name];
        if (s.rows == 0) s.rows = n;
        if (n != s.rows) return;                       // a name keeps the size of its first sample
        s.data.insert(s.data.end(), v, v + n);
    }
    void add(const std::string& name, double v) { add(name, &v, 1); }
    int samples(const std::string& name) const
    {
        auto it = series_.find(name);
        return it == series_.end() || it->second.rows == 0 ? 0 : (int)(it->second.data.size() / it->second.rows);
    }
    // Writes every series to `path` (Level-5 MAT-file, uncompressed, little endian).  Returns false on I/O failure.
    bool flush(const std::string& path) const
    {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        char hdr[128];
        memset(hdr, ' ', 116);
        const char* txt = "MATLAB 5.0 MAT-file, written by qppvm_b200 MatTrace";
        memcpy(hdr, txt, strlen(txt));
        memset(hdr + 116, 0, 8);
        const uint16_t ver = 0x0100, endian = 0x4d49;  // "IM" read as a 16-bit word by a little-endian reader
        memcpy(hdr + 124, &ver, 2); memcpy(hdr + 126, &endian, 2);
        bool ok = fwrite(hdr, 1, 128, f) == 128;
        for (const auto& kv : series_) {
            const Series& s = kv.second;
            if (s.rows == 0) continue;
            const uint32_t cols = (uint32_t)(s.data.size() / s.rows);
            const uint32_t name_len = (uint32_t)kv.first.size(), name_pad = (name_len + 7u) & ~7u;
            const uint32_t data_bytes = (uint32_t)(8u * s.rows * cols);
            const uint32_t total = 16 + 16 + 8 + name_pad + 8 + data_bytes;
            const uint32_t tag[2] = {14u /* miMATRIX */, total};
            const uint32_t flags[4] = {6u /* miUINT32 */, 8u, 6u /* mxDOUBLE_CLASS */, 0u};
            const int32_t dims[4] = {5 /* miINT32 */, 8, (int32_t)s.rows, (int32_t)cols};
            const uint32_t ntag[2] = {1u /* miINT8 */, name_len};
            const uint32_t dtag[2] = {9u /* miDOUBLE */, data_bytes};
            std::vector<char> nm(name_pad, 0);
            memcpy(nm.data(), kv.first.data(), name_len);
            ok = ok && fwrite(tag, 4, 2, f) == 2 && fwrite(flags, 4, 4, f) == 4 && fwrite(dims, 4, 4, f) == 4 &&
                 fwrite(ntag, 4, 2, f) == 2 && (name_pad == 0 || fwrite(nm.data(), 1, name_pad, f) == name_pad) &&
                 fwrite(dtag, 4, 2, f) == 2 && fwrite(s.data.data(), 8, s.data.size(), f) == s.data.size();
        }
        return fclose(f) == 0 && ok;
    }

private:
    struct Series { int rows = 0; std::vector<double> data; };
    std::map<std::string, Series> series_;
};

}  // namespace qppvm
