/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/bits.h>.
 * libosmocore is not installed in this image; this
 * header declares only the names the reference hot-path files use so that they
 * compile unmodified from /root/reference/src.  Written from the API surface, no
 * libosmocore source is vendored. */
#pragma once
#include <stdint.h>
#include <stddef.h>

typedef int8_t  sbit_t;   /* soft bit: -127 (strong 1) .. +127 (strong 0) */
typedef uint8_t ubit_t;   /* unpacked bit, one per byte */
typedef uint8_t pbit_t;   /* packed bits, MSB first */

int osmo_pbit2ubit(ubit_t *out, const pbit_t *in, unsigned int num_bits);
int osmo_ubit2pbit(pbit_t *out, const ubit_t *in, unsigned int num_bits);

static inline unsigned int osmo_pbit_bytesize(unsigned int num_bits)
{
	return (num_bits + 7) / 8;
}

static inline void osmo_store32le(uint32_t x, void *p)
{
	uint8_t *b = (uint8_t *)p;
	b[0] = x & 0xff; b[1] = (x >> 8) & 0xff; b[2] = (x >> 16) & 0xff; b[3] = (x >> 24) & 0xff;
}
