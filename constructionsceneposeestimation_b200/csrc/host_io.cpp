// f3 — label files of a batch written natively ("once kernels are fast, host serialisation
// is the end-to-end limiter at 100 k frames").  The reference opens, dumps and closes one label file per frame
// from Python (gcd.py:608-613, 2071-2072); at a few hundred thousand frames per second the interpreter — not
// the file system — is what limits that loop, so the per-frame open / write / close runs here, without the
// GIL, straight from the D2H text buffer (cspe_format_yolo leaves every frame's text at a fixed stride).
// Plain C++ / POSIX, no CUDA.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "cspe.h"

namespace cspe {
void set_error(const char* fmt, ...);
}

extern "C" int64_t cspe_write_files_host(const char* dir, const char* prefix, int digits, const char* suffix,
                                         int64_t first_id, int count, const char* data_host, int64_t stride,
                                         const int32_t* sizes_host) {
  if (!dir || !prefix || !suffix || digits < 1 || digits > 18 || count < 0 || stride < 0 ||
      (count > 0 && (!data_host || !sizes_host))) {
    cspe::set_error("cspe_write_files_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  std::string path(dir);
  if (!path.empty() && path.back() != '/') path.push_back('/');
  path += prefix;
  const size_t stem = path.size();
  int64_t total = 0;
  char num[32];
  for (int j = 0; j < count; ++j) {
    const int32_t n = sizes_host[j];
    if (n < 0 || n > stride) {
      cspe::set_error("cspe_write_files_host: size %d of file %d outside [0, %lld]", n, j, static_cast<long long>(stride));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    snprintf(num, sizeof(num), "%0*lld", digits, static_cast<long long>(first_id + j));
    path.resize(stem);
    path += num;
    path += suffix;
    const int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
    if (fd < 0) {
      cspe::set_error("cspe_write_files_host: open(%s): %s", path.c_str(), strerror(errno));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    const char* p = data_host + static_cast<int64_t>(j) * stride;
    int64_t left = n;
    while (left > 0) {
      const ssize_t w = write(fd, p, static_cast<size_t>(left));
      if (w < 0) {
        if (errno == EINTR) continue;
        cspe::set_error("cspe_write_files_host: write(%s): %s", path.c_str(), strerror(errno));
        close(fd);
        return CSPE_ERR_INVALID_ARGUMENT;
      }
      p += w;
      left -= w;
    }
    if (close(fd) != 0) {
      cspe::set_error("cspe_write_files_host: close(%s): %s", path.c_str(), strerror(errno));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    total += n;
  }
  return total;
}

extern "C" int64_t cspe_concat_rows_host(const char* data_host, int64_t stride, const int32_t* sizes_host, int count,
                                         char* out_host, int64_t capacity) {
  if (count < 0 || stride < 0 || capacity < 0 || (count > 0 && (!data_host || !sizes_host)) || (capacity > 0 && !out_host)) {
    cspe::set_error("cspe_concat_rows_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  int64_t pos = 0;
  for (int j = 0; j < count; ++j) {
    const int32_t n = sizes_host[j];
    if (n < 0 || n > stride) {
      cspe::set_error("cspe_concat_rows_host: size %d of row %d outside [0, %lld]", n, j, static_cast<long long>(stride));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    if (pos + n > capacity) {
      cspe::set_error("cspe_concat_rows_host: output buffer of %lld bytes is too small", static_cast<long long>(capacity));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    memcpy(out_host + pos, data_host + static_cast<int64_t>(j) * stride, static_cast<size_t>(n));
    pos += n;
  }
  return pos;
}

// The "images" entries of a COCO file for a run of frames (json.dumps default separators):
//   {"id": i, "width": W, "height": H, "file_name": "rgb_%06d.png"}  joined by ", "
// — rgb_%06d.png is the name the capture loop gives its colour images (gcd.py:1672).
extern "C" int64_t cspe_format_coco_images_host(int64_t first_id, int count, int width, int height, char* out_host,
                                                int64_t capacity) {
  if (count < 0 || capacity < 0 || (capacity > 0 && !out_host)) {
    cspe::set_error("cspe_format_coco_images_host: invalid argument");
    return CSPE_ERR_INVALID_ARGUMENT;
  }
  int64_t pos = 0;
  for (int j = 0; j < count; ++j) {
    if (capacity - pos < 160) {
      cspe::set_error("cspe_format_coco_images_host: output buffer of %lld bytes is too small (160 per image)",
                      static_cast<long long>(capacity));
      return CSPE_ERR_INVALID_ARGUMENT;
    }
    const long long id = static_cast<long long>(first_id + j);
    pos += snprintf(out_host + pos, 160, "%s{\"id\": %lld, \"width\": %d, \"height\": %d, \"file_name\": \"rgb_%06lld.png\"}",
                    j ? ", " : "", id, width, height, id);
  }
  return pos;
}
