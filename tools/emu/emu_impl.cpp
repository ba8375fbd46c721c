// Single definition of the emulator's global state + control entry points of libcips3d_b200_emu.so.
// TEST INFRASTRUCTURE ONLY (see c3d_emu.h).
#define C3D_EMU_IMPL
#include "c3d_emu.h"

extern "C" {
// the product loader (cips-3d_b200/_lib.py) refuses any library that exports this symbol
int c3d_emulated(void) { return 1; }
// async_mode: 0 eager, 1 lazy, 2 random; preempt_permille: chance that a primitive yields; sms: reported SM count
void c3d_emu_configure(int async_mode, unsigned long long seed, int preempt_permille, int sms) {
  emu::Config& c = emu::config();
  if (async_mode >= 0) c.async_mode = async_mode;
  c.seed = seed;
  if (preempt_permille >= 0) c.preempt_permille = preempt_permille;
  if (sms > 0) c.sms = sms;
}
}
