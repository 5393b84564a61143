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
uint32_t b200st_crc32c(uint32_t crc, const void* data, size_t n);          /* SSE4.2 crc32 instruction when the CPU has it */
uint32_t b200st_crc32c_table(uint32_t crc, const void* data, size_t n);    /* portable slicing-by-8 path (same result) */
uint32_t b200st_crc32c_mask(uint32_t crc);

/* Walks the records of a whole TFRecord file image buf[0, n).  offsets[i] / lengths[i] receive the position and size of
 * the payload of record i (i < max_records; pass NULL / 0 to only count).  verify: 0 = no checksum, 1 = length
 * checksums, 2 = length and payload checksums.
 * Returns the number of records, or  -1 - (byte position of the damaged record)  on a truncated file or a checksum
 * mismatch (TF raises DataLossError there). */
int64_t b200st_tfrecord_index(const void* buf, size_t n, int64_t* offsets, int64_t* lengths, int64_t max_records, int verify);

/* Writes the 12-byte header and 4-byte footer of one record around a payload: header12 and footer4 are filled in. */
void b200st_tfrecord_frame(const void* payload, size_t n, uint8_t* header12, uint8_t* footer4);

/* ---- tf.train.Example (tensorflow/core/example/{example,feature}.proto: Example{1: Features{1: map<string, Feature{1:
 * BytesList | 2: FloatList | 3: Int64List, each {repeated 1: value}}>}}) — replaces tf.io.parse_single_example with
 * VarLenFeature specs (neurst/data/dataset_utils.py:234-247) for the keys of AudioTFRecordDataset.fields
 * (neurst/data/datasets/audio/audio_dataset.py:296-306).
 * For each of the nkeys NUL-terminated names: kind[i] = 0 (absent or empty feature), 1 bytes, 2 float, 3 int64;
 * off[i] / len[i] = byte range inside rec of the packed values (float: little-endian fp32; int64: varints) or of the FIRST
 * bytes value; count[i] = number of values (-1: the list is stored unpacked / in several chunks — the caller then uses its
 * general decoder).  Returns 0, or -1 for a malformed record. */
#define B200ST_FEATURE_NONE 0
#define B200ST_FEATURE_BYTES 1
#define B200ST_FEATURE_FLOAT 2
#define B200ST_FEATURE_INT64 3
int b200st_example_lookup(const void* rec, size_t n, const char* const* keys, int nkeys, int32_t* kind, int64_t* off, int64_t* len,
                          int64_t* count);
/* Decodes `count` varints (two's-complement int64, as protobuf packs them) from p[0, n); returns the number decoded. */
int64_t b200st_decode_varints(const void* p, size_t n, int64_t* out, int64_t count);

/* Padded batch assembly (tf.data padded_batch with padding value 0 of SpeechToText.create_and_batch_tfds,
 * neurst/tasks/speech2text.py:362-378): row b of dst[B, row_elems] = src[b][0, n_elems[b]) followed by zeros. */
void b200st_pad_rows_f32(float* dst, int64_t row_elems, const float* const* src, const int64_t* n_elems, int32_t B);

#ifdef __cplusplus
}
#endif
#endif
