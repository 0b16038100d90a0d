/* irec_io.h -- C ABI of the index-stream wire format: the integer arithmetic coder and the `.rec` container.
 *
 * the step right after the coding hot path.  The reference's implementation is its only native
 * component (Cython -> C, CPU, sequential per stream): rec/io/entropy_coding.pyx:19-302 (ArithmeticCoder) and
 * rec/io/utils.py:7-216 (write_compressed_code / read_compressed_code).  These entry points are host C++ inside
 * libirec.so; they make no CUDA call and produce byte-identical code strings and files (tests/test_io_parity.py checks
 * them against the reference itself, compiled into oracle/_ref, and against golden files it wrote).
 *
 * Conventions: `int` status return (IREC_OK / IREC_E_* of irec.h, text via irec_last_error_string()); all buffers are
 * caller-owned host memory; a code string is one byte per bit (0 or 1), the order the reference's list of '0'/'1'
 * characters has.
 */
#ifndef IREC_IO_H
#define IREC_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ArithmeticCoder.encode (rec/io/entropy_coding.pyx:51-117).  counts[n_symbols] = P (unnormalised masses, all > 0),
 * message[n_message] symbols in [0, n_symbols).  Writes at most `capacity` bits; *out_n_bits is the full code length
 * (IREC_E_CAPACITY if it did not fit: call again with a larger buffer). */
int irec_ac_encode(const int64_t* counts, int n_symbols, int precision, const int64_t* message, int64_t n_message,
                   uint8_t* out_bits, int64_t capacity, int64_t* out_n_bits);

/* ArithmeticCoder.decode / decode_fast (rec/io/entropy_coding.pyx:121-208, 212-302): decodes until symbol 0 (the
 * end-of-message symbol, included in the output).  Symbol lookup is a binary search over the cumulative masses -- the
 * result of the reference's linear scan (decode) and of its AVL interval tree (decode_fast, rec/io/data_structures.py:
 * 184-210).  IREC_E_INVALID for a code that does not decode (the reference loops forever / raises on such input). */
int irec_ac_decode(const int64_t* counts, int n_symbols, int precision, const uint8_t* bits, int64_t n_bits,
                   int64_t* out_message, int64_t capacity, int64_t* out_n_message);

/* Description of one coded image for the container: n_res_blocks latent tensors ("residual blocks"), tensor r split
 * into num_blocks[r] coder-blocks; num_aux (concatenated over tensors, then blocks) = number of indices of each
 * coder-block; indices = all indices in that order. */
typedef struct {
    uint32_t seed, block_size, max_index;
    uint32_t image_h, image_w;
    uint16_t image_c;
    uint16_t uses_num_aux_counts_file, uses_index_counts_file;   /* header flags (rec/io/utils.py:88-89) */
    int32_t n_res_blocks;
} irec_rec_header_t;

/* write_compressed_code (rec/io/utils.py:7-106) into a byte buffer.  index_counts: NULL for the reference's default
 * masses (1 for the end symbol, 1001 for every index, rec/io/utils.py:31-35) or max_index + 1 masses (the contents of
 * the reference's index_counts_file).  The number-of-auxiliary-variable streams always use the default masses (1, 101,
 * ... over max + 2 symbols, rec/io/utils.py:43-49).  *out_bytes = size of the file image. */
int irec_rec_pack(const irec_rec_header_t* header, const int32_t* num_blocks, const int32_t* num_aux, const int64_t* indices,
                  const int64_t* index_counts, uint8_t* out, int64_t capacity, int64_t* out_bytes);

/* read_compressed_code (rec/io/utils.py:109-216), step 1: the static header (28 bytes). */
int irec_rec_read_header(const uint8_t* file, int64_t file_bytes, irec_rec_header_t* header);

/* step 2: decode everything.  num_blocks[n_res_blocks]; num_aux / indices are filled up to their capacities and the
 * required element counts are returned (IREC_E_CAPACITY when a buffer was too small: call again). */
int irec_rec_unpack(const uint8_t* file, int64_t file_bytes, const int64_t* index_counts, int32_t* num_blocks,
                    int32_t* num_aux, int64_t num_aux_capacity, int64_t* out_n_num_aux,
                    int64_t* indices, int64_t indices_capacity, int64_t* out_n_indices);

/* the same through the file system (fopen/fwrite/fread) */
int irec_rec_write_file(const char* path, const irec_rec_header_t* header, const int32_t* num_blocks, const int32_t* num_aux,
                        const int64_t* indices, const int64_t* index_counts, int64_t* out_bytes);

#ifdef __cplusplus
}
#endif
#endif
