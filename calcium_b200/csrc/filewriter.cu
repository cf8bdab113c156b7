// filewriter.cu -- host-side writer for the files K3 leaves in (pinned) host memory.
//
// The batch path (reference main.py:294-341: save_image + np.savez_compressed per frame) ends in tens of
// thousands of small files per call: 148 medium pouches x 90 frames x (.png + .npz) = 26 640 files, 0.55 GB.
// Writing them from Python costs more than producing them on the GPU (one interpreter-level open / write /
// close and one bytes() copy per file).  cab200_write_files takes the whole chunk at once -- paths, offsets and
// sizes into the arena -- and writes it with a pool of native threads straight from the arena.  No device code.
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace cab200;

extern "C" int cab200_write_files(int64_t n_files, const char *paths, const int64_t *path_off, const uint8_t *base,
                                  const int64_t *offset, const int64_t *size, int n_threads, int64_t *bytes_out)
{
    if (n_files < 0 || (n_files > 0 && (!paths || !path_off || !base || !offset || !size))) {
        set_error("cab200_write_files: bad argument");
        return CAB200_EINVAL;
    }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if ((int64_t)n_threads > n_files) n_threads = (int)(n_files > 0 ? n_files : 1);
    std::atomic<int64_t> next(0), total(0);
    std::mutex err_mu;
    std::string first_err;
    auto work = [&]() {
        int64_t done = 0;
        for (;;) {
            const int64_t lo = next.fetch_add(16);             // 16 files per grab
            if (lo >= n_files) break;
            const int64_t hi = lo + 16 < n_files ? lo + 16 : n_files;
            for (int64_t i = lo; i < hi; ++i) {
                const char *path = paths + path_off[i];
                const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
                if (fd < 0) {
                    std::lock_guard<std::mutex> lk(err_mu);
                    if (first_err.empty()) first_err = std::string(path) + ": " + strerror(errno);
                    continue;
                }
                const uint8_t *p = base + offset[i];
                int64_t left = size[i];
                while (left > 0) {
                    const ssize_t w = write(fd, p, (size_t)left);
                    if (w < 0) {
                        if (errno == EINTR) continue;
                        std::lock_guard<std::mutex> lk(err_mu);
                        if (first_err.empty()) first_err = std::string(path) + ": " + strerror(errno);
                        break;
                    }
                    p += w; left -= w;
                }
                close(fd);
                done += size[i] - left;
            }
        }
        total += done;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (bytes_out) *bytes_out = total.load();
    if (!first_err.empty()) {
        set_error("cab200_write_files: %s", first_err.c_str());
        return CAB200_EINVAL;
    }
    return CAB200_OK;
}
