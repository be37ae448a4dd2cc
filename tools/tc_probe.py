"""Diagnostic runner for the tcgen05 kernel: prints one line per case and never stops at the first failure.
A device-side trap poisons the CUDA context, so on any CUDA error the remaining cases re-run one per process."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    import torch
    from tc_cases import HALO_CASES, TC_CASES, run_tc_case
    if os.environ.get("PROBE_SET", "tc") == "halo":
        TC_CASES = HALO_CASES
    variant = int(os.environ.get("PROBE_VARIANT", "0"))
    idx = [int(a) for a in sys.argv[1:]] or list(range(len(TC_CASES)))
    single = len(sys.argv) > 1
    for i in idx:
        c = TC_CASES[i]
        for dt in (torch.float16, torch.bfloat16):
            try:
                rel, mx, bad = run_tc_case(c, dt, variant=variant)
                print(f"[{i:2d}] {c[0]:28s} {str(dt)[6:]:9s} rel {rel:.3e} max {mx:.3e} bad_rows {bad:.4f} "
                      f"{'OK' if bad == 0 and rel < 1e-2 else 'WRONG'}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"[{i:2d}] {c[0]:28s} {str(dt)[6:]:9s} ERROR {str(e)[:200]}", flush=True)
                if single:
                    return 1
                for j in idx[idx.index(i) + 1:]:
                    subprocess.run([sys.executable, __file__, str(j)], timeout=120)
                return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
