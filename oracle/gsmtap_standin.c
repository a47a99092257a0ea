/* TEST INFRASTRUCTURE ONLY - stand-in for the reference's tetra_gsmtap.c (src/tetra_gsmtap.c:31-82), which
 * needs libosmocore's gsmtap headers and socket helpers (absent here).  GSMTAP is a UDP side channel to
 * wireshark; nothing on stdout depends on it.  Kept: the one side effect the rest of the program reads,
 * tms->tsn = ts (tetra_gsmtap.c:52, used in the traffic dump file name, tetra_lower_mac.c:205). */
#include <stdint.h>
#include <stddef.h>

#include <osmocom/core/msgb.h>
#include <osmocom/core/bits.h>

#include "tetra_common.h"
#include "tetra_tdma.h"

struct msgb *tetra_gsmtap_makemsg(struct tetra_tdma_time *tm, enum tetra_log_chan lchan, uint8_t ts, uint8_t ss,
				  int8_t signal_dbm, uint8_t snr, const ubit_t *bitdata, unsigned int bitlen,
				  struct tetra_mac_state *tms)
{
	(void)tm; (void)lchan; (void)ss; (void)signal_dbm; (void)snr; (void)bitdata; (void)bitlen;
	tms->tsn = ts;
	return NULL;
}

int tetra_gsmtap_sendmsg(struct msgb *msg) { (void)msg; return 0; }
int tetra_gsmtap_init(const char *host, uint16_t port) { (void)host; (void)port; return 0; }
