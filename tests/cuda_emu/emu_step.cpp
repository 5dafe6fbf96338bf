// Runs the CUDA step kernels through the CPU thread emulator on host arrays
// that are already in the device (batch-innermost) layout.
#include "cuda_emu.h"

#include "../../opensc2_b200/csrc/step_launch.h"

extern "C" int osc2_emu_step(const osc2_topology *tp, const StepDev *p, int fast, int split)
{
    EmuLauncher L;
    StepDev q = *p;
    std::vector<int> leaves, prog;
    osc2_build_norm_plan(0, q.N, leaves, prog);
    std::vector<double> part(leaves.size() / 2 * 3 * (size_t)q.d * q.B, std::nan(""));
    q.norm_leaf = leaves.data(); q.norm_prog = prog.data(); q.norm_part = part.data();
    q.n_leaves = (int)(leaves.size() / 2); q.n_prog = (int)prog.size();
    (void)split;
    q.cpw = fast ? osc2_fast_cpw(q.d) : 1;
    q.NT = fast ? (int)osc2_fast_tiles(q.d, q.B) : q.B;
    std::vector<unsigned char> smap;
    q.s_slots = osc2_build_smap(*tp, smap);
    q.smap = smap.data();
    std::vector<unsigned short> sinv((size_t)(q.s_slots > 0 ? q.s_slots : 1), 0);
    for (size_t i = 0; i < smap.size(); ++i) if (smap[i]) sinv[smap[i]] = (unsigned short)((i % q.d) * 32 + (i / q.d) * q.cpw);
    q.sinv = sinv.data();
    q.gm_prezeroed = fast ? 1 : 0;      // emu.py hands over a zeroed `gm` on the fast path, as osc2_create does
    osc2_run_step(L, tp, q, fast != 0, tp->theta);
    return (int)L.launches;
}
extern "C" int osc2_emu_has_fast(int d) { return osc2_has_fast_path(d) ? 1 : 0; }
extern "C" int osc2_emu_gm_slots(int d, int M, int S, int fast) { return osc2_gm_slots(d, M, S, fast != 0); }
extern "C" long long osc2_emu_gm_doubles(int d, int M, int S, int fast, long long E, long long B) { return (long long)osc2_gm_doubles(d, M, S, fast != 0, (size_t)E, (size_t)B); }
extern "C" long long osc2_emu_spill_doubles(int d, int fast, long long N, long long B) { return (long long)osc2_spill_doubles(d, fast != 0, (size_t)N, (size_t)B); }

extern "C" int osc2_emu_transpose(const double *in, double *out, long long rows, long long cols)
{
    EmuLauncher L;
    const long long tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
    L.launch(0, osc2_transpose_kernel, (unsigned)tiles, 256u, (size_t)0, in, out, rows, cols);
    return 0;
}

extern "C" int osc2_emu_sizeof_stepdev() { return (int)sizeof(StepDev); }
extern "C" int osc2_emu_sizeof_topology() { return (int)sizeof(osc2_topology); }
