// Precomputed-flow ingestion: the ".flo" reader of the reference's file mode (-f <flowdir>), host side.
// Replaces ReadFlowFile (src/stabilization/flowIO.cpp:31-78) and the file naming of
// FileStabilizer::retrieveOpticalFlow (src/stabilization/stabilizefiles.cpp:135-149).
//
// Format (flowIO.cpp:5-19): float tag 202021.25 ("PIEH"), int32 width, int32 height, then width*height
// interleaved little-endian (u, v) float pairs in row order.  The reference reads row by row into a
// std::vector and later copies that pageable buffer to the device; here the payload is read with one fread
// straight into the caller's buffer -- in the pipeline object a pinned landing buffer, so the H2D copy that
// follows is asynchronous.  Every reference exception has its own error code whose vsc_error_string() is the
// reference's message.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <thread>

#include "vsc_common.cuh"

namespace vsc {

static int flo_open(const char* path, FILE** fp, int* width, int* height)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return VSC_E_FLO_OPEN;                                   // flowIO.cpp:34-36
    float tag = 0.0f;
    int32_t w = 0, h = 0;
    if (std::fread(&tag, sizeof(float), 1, f) != 1 || std::fread(&w, sizeof(int32_t), 1, f) != 1
        || std::fread(&h, sizeof(int32_t), 1, f) != 1) {
        std::fclose(f);
        return VSC_E_FLO_HEADER;                                 // :41-45
    }
    if (tag != 202021.25f) {
        std::fclose(f);
        return VSC_E_FLO_TAG;                                    // :47-49
    }
    *width = w;   // reported even when rejected below, like the reference's by-reference outputs
    *height = h;
    if (w < 1 || w > 99999) {
        std::fclose(f);
        return VSC_E_FLO_WIDTH;                                  // :52-54
    }
    if (h < 1 || h > 99999) {
        std::fclose(f);
        return VSC_E_FLO_HEIGHT;                                 // :56-58
    }
    *fp = f;
    return VSC_OK;
}

int flo_read_into(const char* path, float* dst, size_t cap_floats, int* width, int* height)
{
    FILE* f = nullptr;
    const int rc = flo_open(path, &f, width, height);
    if (rc)
        return rc;
    const size_t n = static_cast<size_t>(*width) * static_cast<size_t>(*height) * 2;
    if (n > cap_floats) {
        std::fclose(f);
        return VSC_E_WORKSPACE;
    }
    // Large payloads of regular files (a 4K flow is 66 MB) are copied out of the page cache by a few threads
    // with pread, each its own byte range: a single fread is bound by one core's copy bandwidth.
    struct stat st;
    const size_t bytes = n * sizeof(float);
    const int fd = fileno(f);
    if (bytes >= (8u << 20) && fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
        const size_t have = static_cast<size_t>(st.st_size);
        if (have != 12 + bytes) {
            std::fclose(f);
            return have < 12 + bytes ? VSC_E_FLO_SHORT : VSC_E_FLO_LONG;   // :67-69, :72-74
        }
        // two files of a frame are read concurrently (stabilizer.cu): half the cores each, at most 8
        unsigned nt = std::thread::hardware_concurrency() / 2;
        nt = nt > 8 ? 8 : (nt < 1 ? 1 : nt);
        const size_t part = (bytes / nt + 4095) & ~static_cast<size_t>(4095);
        bool ok[8] = {true, true, true, true, true, true, true, true};
        auto work = [&](unsigned k) {
            size_t off = k * part;
            const size_t end = off + part < bytes ? off + part : bytes;
            char* out = reinterpret_cast<char*>(dst);
            while (off < end) {
                const ssize_t got = pread(fd, out + off, end - off, static_cast<off_t>(12 + off));
                if (got <= 0) {
                    ok[k] = false;
                    return;
                }
                off += static_cast<size_t>(got);
            }
        };
        std::thread th[7];
        for (unsigned k = 1; k < nt; ++k)
            th[k - 1] = std::thread(work, k);
        work(0);
        for (unsigned k = 1; k < nt; ++k)
            th[k - 1].join();
        std::fclose(f);
        for (unsigned k = 0; k < nt; ++k)
            if (!ok[k])
                return VSC_E_FLO_SHORT;
        return VSC_OK;
    }
    if (std::fread(dst, sizeof(float), n, f) != n) {
        std::fclose(f);
        return VSC_E_FLO_SHORT;                                  // :67-69
    }
    if (std::fgetc(f) != EOF) {
        std::fclose(f);
        return VSC_E_FLO_LONG;                                   // :72-74
    }
    std::fclose(f);
    return VSC_OK;
}

}  // namespace vsc

extern "C" int vsc_flo_read_header(const char* path, int* width, int* height)
{
    if (!path || !width || !height)
        return VSC_E_INVALID;
    FILE* f = nullptr;
    const int rc = vsc::flo_open(path, &f, width, height);
    if (f)
        std::fclose(f);
    return rc;
}

extern "C" int vsc_flo_read(const char* path, float* dst, size_t dst_capacity_floats, int* width, int* height)
{
    if (!path || !dst || !width || !height)
        return VSC_E_INVALID;
    return vsc::flo_read_into(path, dst, dst_capacity_floats, width, height);
}

extern "C" int vsc_flo_frame_path(const char* flow_dir, int frame, int backward, char* out, size_t out_capacity)
{
    if (!flow_dir || !out || out_capacity == 0)
        return VSC_E_INVALID;
    const size_t len = std::strlen(flow_dir);
    const char* sep = (len > 0 && flow_dir[len - 1] != '/') ? "/" : "";
    // formatIndex (stabilizefiles.cpp:99-101): decimal, zero-padded to at least 6 digits
    const int n = std::snprintf(out, out_capacity, "%s%sframe_%06d%s.flo", flow_dir, sep, frame, backward ? "_bwd" : "");
    return (n < 0 || static_cast<size_t>(n) >= out_capacity) ? VSC_E_INVALID : VSC_OK;
}
