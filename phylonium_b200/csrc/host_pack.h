// 2-bit packing of sequences on the host (host_pack.cpp) for the packed upload (staging.h).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace phy
{

// Packs n bytes over {A,C,G,T,!} into (n + 3) / 4 bytes: base k of a group of four in bits
// 2k, 2k + 1, code (c >> 1) & 3 (A 0, C 1, T 2, G 3); '!' packs as 0 and its position is
// appended to bangs (at most cap entries are written; *nbangs counts all of them).
// Returns 1 if a byte outside the alphabet was seen, else 0.
int pack_2bit(const uint8_t *src, size_t n, uint8_t *dst, uint32_t *bangs, uint32_t cap, uint32_t *nbangs);

} // namespace phy
