/* TEST INFRASTRUCTURE (oracle/_ref build only).
 * Declarations of the libfor entry points the reference calls (SURVEY.md §2.1 call sites:
 * src/array_base.cpp:14; src/array.cpp:4,8,13,30,51,108; src/sorted_array.cpp:12,30,59,89,98,108,126-141;
 * include/sorted_array.h:20; include/array.h:17).
 * libfor (github.com/cruppstahl/libfor @ 49611808d08d4e47116aa2a3ddcabeb418f405f7, cmake/For.cmake:3) is a
 * build-time download of the reference and is NOT present on this machine; oracle/libfor_port.c restates its
 * published format: [u32 base][u8 bits][values-base packed LSB-first, `bits` bits each]. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
uint32_t for_compressed_size_bits(uint32_t length, uint32_t bits);
uint32_t for_compress_sorted(const uint32_t* in, uint8_t* out, uint32_t length);
uint32_t for_compress_unsorted(const uint32_t* in, uint8_t* out, uint32_t length);
uint32_t for_uncompress(const uint8_t* in, uint32_t* out, uint32_t length);
uint32_t for_append_sorted(uint8_t* in, uint32_t length, uint32_t value);
uint32_t for_append_unsorted(uint8_t* in, uint32_t length, uint32_t value);
uint32_t for_select(const uint8_t* in, uint32_t index);
uint32_t for_select_bits(const uint8_t* in, uint32_t base, uint32_t bits, uint32_t index);
uint32_t for_linear_search(const uint8_t* in, uint32_t length, uint32_t value);
uint32_t for_lower_bound_search(const uint8_t* in, uint32_t length, uint32_t value, uint32_t* actual);
#ifdef __cplusplus
}
#endif
