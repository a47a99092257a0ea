/* host-side bit packing rate (one 0/1 byte per bit -> eight bits per byte) with T threads: is the CPU a faster road to the
 * GPU than PCIe for the reference's byte-per-bit format?  gcc -O3 -mavx2 -pthread host_pack.c -o host_pack */
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
static const uint8_t *src; static uint32_t *dst; static size_t n_bytes; static int T;
static void *work(void *arg)
{
	const size_t t = (size_t)arg, per = (n_bytes / 32 / T) * 32, lo = t * per, hi = t + 1 == (size_t)T ? n_bytes : lo + per;
	for (size_t i = lo; i + 32 <= hi; i += 32) {
		const __m256i v = _mm256_loadu_si256((const __m256i *)(src + i));
		dst[i >> 5] = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 7));
	}
	return NULL;
}
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int main(int argc, char **argv)
{
	n_bytes = (size_t)(argc > 1 ? atof(argv[1]) : 4e9);
	uint8_t *s = malloc(n_bytes); dst = malloc(n_bytes / 8 + 64);
	for (size_t i = 0; i < n_bytes; i++) s[i] = (uint8_t)((i * 2654435761u) >> 31 & 1);
	memset(dst, 0, n_bytes / 8 + 64);
	src = s;
	printf("online cpus %ld\n", sysconf(_SC_NPROCESSORS_ONLN));
	int ts[] = {1, 2, 4, 8, 12, 16, 24, 32, 48, 64};
	for (unsigned k = 0; k < sizeof ts / sizeof *ts; k++) {
		T = ts[k];
		if (T > 2 * sysconf(_SC_NPROCESSORS_ONLN)) break;
		pthread_t th[64];
		double best = 1e9;
		for (int rep = 0; rep < 3; rep++) {
			const double t0 = now();
			for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, work, (void *)(size_t)t);
			for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
			const double dt = now() - t0; if (dt < best) best = dt;
		}
		printf("threads %2d: %.1f GB/s of byte-per-bit input\n", T, n_bytes / best / 1e9);
	}
	return 0;
}
