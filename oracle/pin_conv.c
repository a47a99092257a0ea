/* TEST INFRASTRUCTURE ONLY - `make -C oracle pin-libosmocore`.
 *
 * Pins the one part of the oracle that is restated from memory: libosmocore's osmo_conv_decode on NOISY input
 * (lower_mac/viterbi_cch.c:58-66 calls it; the library is absent from /root/reference and from the build image).
 * Where a libosmocore development package is installed this program links the REAL osmo_conv_decode next to the
 * two restatements of oracle/osmo_standin.c, decodes the same deterministic noisy blocks (the three block lengths
 * of the receive chain, BER 0 .. 15 %, hard +-127 inputs with the 2/3 puncturing's erasures) with all of them
 * and reports which setting of include/tetra_tie_rule.h reproduces the library bit for bit.
 *
 *   exit 0  the compiled-in TETRA_VITERBI_TIE_DEFAULT matches the library on every block
 *   exit 1  it does not (the message names the setting that does, if any): flip TETRA_VITERBI_TIE_DEFAULT
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <osmocom/core/bits.h>
#include <osmocom/core/conv.h>

#include "../include/tetra_tie_rule.h"

int oracle_tetra_cch_decode(int variant, const sbit_t *in, int n, ubit_t *out);
void oracle_conv_set_tie(int tie);

static uint64_t rng_state = 0x7E7A7E7Aull;
static uint32_t rnd(void)
{
	rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
	return (uint32_t)(rng_state >> 16);
}

/* mother code of tetra_conv_enc.c:43-74 == tables viterbi_cch.c:35-48, rebuilt from the polynomials */
static uint8_t nout[16][2], nst[16][2];

static void build_tables(void)
{
	for (unsigned s = 0; s < 16; s++)
		for (unsigned b = 0; b < 2; b++) {
			unsigned d1 = s & 1, d2 = (s >> 1) & 1, d3 = (s >> 2) & 1, d4 = (s >> 3) & 1;
			unsigned g1 = b ^ d1 ^ d4, g2 = b ^ d2 ^ d3 ^ d4, g3 = b ^ d1 ^ d2 ^ d4, g4 = b ^ d1 ^ d3 ^ d4;
			nout[s][b] = (g1 << 3) | (g2 << 2) | (g3 << 1) | g4;
			nst[s][b] = ((s << 1) | b) & 15;
		}
}

int main(void)
{
	static const int lens[3] = { 80, 144, 288 };
	static const int ber_per_1000[6] = { 0, 10, 30, 60, 100, 150 };
	long blocks = 0, bad[2] = { 0, 0 };
	build_tables();
	for (int li = 0; li < 3; li++)
		for (int bi = 0; bi < 6; bi++)
			for (int rep = 0; rep < 200; rep++) {
				const int n = lens[li];
				uint8_t t2[292];
				sbit_t soft[4 * 292];
				ubit_t real[288], a[288], g[288];
				memset(t2, 0, sizeof(t2));
				for (int i = 0; i < n - 4; i++) t2[i] = rnd() & 1;
				memset(soft, 0, sizeof(soft));
				unsigned st = 0;
				for (int t = 0; t < n; t++) {
					const unsigned o = nout[st][t2[t]];
					st = nst[st][t2[t]];
					/* rate 2/3: G1,G2 on even steps, G1 on odd steps survive (tetra_conv_enc.c:96,128-134) */
					const int keep = (t & 1) ? 1 : 2;
					for (int j = 0; j < keep; j++) {
						int bit = (o >> (3 - j)) & 1;
						if ((int)(rnd() % 1000) < ber_per_1000[bi]) bit ^= 1;
						soft[4 * t + j] = bit ? -127 : 127;          /* viterbi.c:13-22 */
					}
				}
				struct osmo_conv_code code;
				memset(&code, 0, sizeof(code));
				code.N = 4; code.K = 5; code.len = n;
				code.next_output = nout; code.next_state = nst;
				oracle_conv_set_tie(TETRA_VITERBI_TIE_DEFAULT);      /* (only matters in the self-check build, where the "library" is the stand-in) */
				osmo_conv_decode(&code, soft, real);
				blocks++;
				for (int tie = 0; tie < 2; tie++) {
					oracle_conv_set_tie(tie);
					oracle_tetra_cch_decode(0, soft, n, a);
					oracle_tetra_cch_decode(1, soft, n, g);
					if (memcmp(a, real, n) || memcmp(g, real, n)) bad[tie]++;
				}
			}
	printf("pin-libosmocore: %ld noisy blocks; restatement differs from the library on %ld (ties keep s>>1) / %ld (ties keep (s>>1)|8)\n",
	       blocks, bad[0], bad[1]);
	if (bad[TETRA_VITERBI_TIE_DEFAULT] == 0) {
		printf("pin-libosmocore: TETRA_VITERBI_TIE_DEFAULT=%d reproduces libosmocore bit for bit - Viterbi parity PINNED\n",
		       TETRA_VITERBI_TIE_DEFAULT);
		return 0;
	}
	if (bad[1 - TETRA_VITERBI_TIE_DEFAULT] == 0)
		printf("pin-libosmocore: set TETRA_VITERBI_TIE_DEFAULT=%d in include/tetra_tie_rule.h and rebuild\n", 1 - TETRA_VITERBI_TIE_DEFAULT);
	else
		printf("pin-libosmocore: NEITHER setting reproduces this libosmocore - the restatement needs work\n");
	return 1;
}
