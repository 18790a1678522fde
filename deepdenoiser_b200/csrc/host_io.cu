// Host-side I/O helpers of the training input pipeline (no device code): CRC-32C (Castagnoli) as used by the TFRecord
// framing that TFRecordsCreator.py:221-230 writes through tf.python_io.TFRecordWriter.  Slicing-by-8, table built once.
#include <stddef.h>
#include <stdint.h>

#include "../../include/dd_b200.h"

namespace {
struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFFu];
  }
};
const Crc32cTables& tables() {
  static const Crc32cTables k;
  return k;
}
}  // namespace

extern "C" uint32_t dd_crc32c(const void* data, size_t size) {
  const Crc32cTables& T = tables();
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t crc = 0xFFFFFFFFu;
  while (size >= 8) {
    const uint32_t lo = crc ^ (static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) |
                               (static_cast<uint32_t>(p[2]) << 16) | (static_cast<uint32_t>(p[3]) << 24));
    crc = T.t[7][lo & 0xFFu] ^ T.t[6][(lo >> 8) & 0xFFu] ^ T.t[5][(lo >> 16) & 0xFFu] ^ T.t[4][lo >> 24] ^
          T.t[3][p[4]] ^ T.t[2][p[5]] ^ T.t[1][p[6]] ^ T.t[0][p[7]];
    p += 8; size -= 8;
  }
  while (size--) crc = (crc >> 8) ^ T.t[0][(crc ^ *p++) & 0xFFu];
  return crc ^ 0xFFFFFFFFu;
}
