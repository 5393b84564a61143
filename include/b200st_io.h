/* libb200st_io — host-side helpers of the input edge of the SpeechTransformer path (plain C, no CUDA):
 * the record framing of the TFRecord files the reference trains from.
 *
 * Replaces, for this path, what the reference gets from TensorFlow:
 *   tf.data.TFRecordDataset(f, buffer_size=...)           neurst/data/dataset_utils.py:317-320
 *   tf.io.TFRecordWriter / tf.io.tf_record_iterator       neurst/data/datasets/audio/audio_dataset.py:214-237,
 *                                                         neurst/utils/misc.py (take_one_record)
 * Format (tensorflow/core/lib/io/record_writer.h; third-party, absent from the reference tree — restated from its
 * published description and pinned on the reference's own fixtures tests/examples/train.tfrecords-0000?-of-00004):
 *   uint64 length (LE) | uint32 masked_crc32c(length) | byte data[length] | uint32 masked_crc32c(data)
 *   masked(c) = ((c >> 15) | (c << 17)) + 0xa282ead8,  crc32c = CRC-32/Castagnoli (reflected poly 0x82F63B78)
 */
#ifndef B200ST_IO_H_
#define B200ST_IO_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int b200st_io_version(void);

/* CRC-32C of data[0, n) continuing from `crc` (0 for a fresh checksum). */
uint32_t b200st_crc32c(uint32_t crc, const void* data, size_t n);
uint32_t b200st_crc32c_mask(uint32_t crc);

/* Walks the records of a whole TFRecord file image buf[0, n).  offsets[i] / lengths[i] receive the position and size of
 * the payload of record i (i < max_records; pass NULL / 0 to only count).  verify: 0 = no checksum, 1 = length
 * checksums, 2 = length and payload checksums.
 * Returns the number of records, or  -1 - (byte position of the damaged record)  on a truncated file or a checksum
 * mismatch (TF raises DataLossError there). */
int64_t b200st_tfrecord_index(const void* buf, size_t n, int64_t* offsets, int64_t* lengths, int64_t max_records, int verify);

/* Writes the 12-byte header and 4-byte footer of one record around a payload: header12 and footer4 are filled in. */
void b200st_tfrecord_frame(const void* payload, size_t n, uint8_t* header12, uint8_t* footer4);

#ifdef __cplusplus
}
#endif
#endif
