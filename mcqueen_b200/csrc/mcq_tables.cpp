// mcq_tables.cpp - codon substitution classes derived from the standard genetic code.
//
// The reference stores four 64x64 0/1 matrices NS_TS, NS_TV, S_TS, S_TV and their row sums
// alpha_S, alpha_V, beta_S, beta_V as literals (constants.c:405-688).  They are a pure function
// of the genetic code in TCAG order (amino_acids.c:20-28): two codons that differ at exactly one
// position are classed by (non-)synonymous x transition (T<->C, A<->G) / transversion; stop
// codons are treated as one amino-acid class '*'.  tests/test_tables.py checks this derivation
// against the reference's compiled tables.
#include <stdint.h>

static const char kAmino[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

// 0 none/multi-step, 1 NS_TS, 2 NS_TV, 3 S_TS, 4 S_TV
extern "C" int mcq_codon_class(int c1, int c2) {
  int ndiff = 0, a = 0, b = 0;
  for (int p = 0; p < 3; p++) {
    const int x = (c1 >> (4 - 2 * p)) & 3, y = (c2 >> (4 - 2 * p)) & 3;
    if (x != y) { ndiff++; a = x; b = y; }
  }
  if (ndiff != 1) return 0;
  const bool ts = ((a ^ b) == 1);
  const bool syn = kAmino[c1] == kAmino[c2];
  return syn ? (ts ? 3 : 4) : (ts ? 1 : 2);
}

extern "C" void mcq_host_tables(uint8_t cls[64 * 64], uint8_t cnt[4][64]) {
  for (int c1 = 0; c1 < 64; c1++) {
    for (int k = 0; k < 4; k++) cnt[k][c1] = 0;
    for (int c2 = 0; c2 < 64; c2++) {
      const int c = mcq_codon_class(c1, c2);
      cls[c1 * 64 + c2] = (uint8_t)c;
      // cnt order: alpha_S (NS_TS), alpha_V (NS_TV), beta_S (S_TS), beta_V (S_TV)
      if (c) cnt[c - 1][c1]++;
    }
  }
}
