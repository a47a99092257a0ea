/* TEST INFRASTRUCTURE ONLY - stand-in for libosmocore <osmocom/core/gsmtap_util.h>: the three calls
 * tetra_gsmtap.c makes.  gsmtap_glue.c implements them: the "socket" is a capture buffer. */
#pragma once
#include <stdint.h>

struct msgb;
struct gsmtap_inst;

struct gsmtap_inst *gsmtap_source_init(const char *host, uint16_t port, int ofd_wq_mode);
int gsmtap_source_add_sink(struct gsmtap_inst *gti);
int gsmtap_sendmsg(struct gsmtap_inst *gti, struct msgb *msg);
