/*
 * tetra_tie_rule.h - the ONE compile-time switch for the Viterbi tie rule, shared by the CUDA kernels
 * (osmo-tetra_b200/csrc) and the CPU oracle (oracle/tetra_oracle.c, oracle/osmo_standin.c).
 *
 * The reference decodes with libosmocore's osmo_conv_decode (lower_mac/viterbi_cch.c:58-66).  libosmocore is
 * neither in /root/reference nor in this image, so its behaviour on EQUAL path metrics is restated from the
 * published algorithm (conv_acc_generic.c "sum0 >= sum1" / conv.c strict "<" while scanning states upwards):
 * the survivor is the predecessor whose oldest register bit is 0, i.e. state s>>1 rather than (s>>1)|8 in the
 * numbering of viterbi_cch.c:42-48.  No test of the reference pins this.
 *
 *   TETRA_VITERBI_TIE_DEFAULT 0   ties keep predecessor  s>>1      (libosmocore as restated; the default)
 *   TETRA_VITERBI_TIE_DEFAULT 1   ties keep predecessor (s>>1)|8   (the other rule)
 *
 * A maintainer with a libosmocore build runs `make -C oracle pin-libosmocore`; should it report the other rule,
 * flipping this one define (or -DTETRA_VITERBI_TIE_DEFAULT=1) changes kernels and oracle together.  Both
 * settings are also selectable at run time (tb200_options.viterbi_tie, orc_set_tie) and both are under test.
 */
#ifndef TETRA_TIE_RULE_H
#define TETRA_TIE_RULE_H

#define TETRA_TIE_KEEPS_LOW_PRED   0
#define TETRA_TIE_KEEPS_HIGH_PRED  1

#ifndef TETRA_VITERBI_TIE_DEFAULT
#define TETRA_VITERBI_TIE_DEFAULT TETRA_TIE_KEEPS_LOW_PRED
#endif

#endif
