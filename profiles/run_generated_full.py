"""Full-size run of the GENERATED programs on the B200: the reference's own CUDA kernels (stock CUDAGenerator,
sm_100a) vs the retargeted generator (libgala_b200 bindings, device-side data prep, fused transforms).
Binaries from host/codegen/build_models.sh; the dataset is synthesised here in the on-disk .npy format.

    python profiles/run_generated_full.py <Reddit|Products> <program> [<program> ...]

Per program and generator: mean forward ms and forward+backward+Adam ms as the program prints them
(common.h:1571-1587), start-up seconds = wall clock of the process minus its 100 epochs (data load, format
construction, H2D, CUDA context), and the epoch-1 CHECK line (checksum, loss) for parity."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CG = os.path.join(ROOT, "gala-gnn-acceleration-language_b200", "host", "codegen")
SHAPES = {"Reddit": ("232965", "114615892", "602", "41"), "Products": ("2449029", "123718280", "100", "47")}


def run(ds, programs, kinds=("ref", "b200"), keep_data=False, log=None):
    """Synthesises the dataset, runs every program with every generator, returns one dict per run."""
    data = os.path.join(CG, "_models", "Data", ds)
    subprocess.run(["rm", "-rf", data])
    subprocess.run([sys.executable, os.path.join(CG, "make_npy_dataset.py"), data, *SHAPES[ds]], check=True,
                   stdout=subprocess.DEVNULL if log is None else log)
    out = []
    for prog in programs:
        for kind in kinds:
            cwd = os.path.join(CG, "_models", f"{prog}_{kind}", "build")
            rec = {"program": prog, "generator": kind, "dataset": ds}
            if not os.path.exists(os.path.join(cwd, "gala_model")):
                rec["error"] = "binary missing (host/codegen/build_models.sh)"
                out.append(rec)
                continue
            t0 = time.perf_counter()
            r = subprocess.run(["./gala_model"], cwd=cwd, capture_output=True, text=True)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                rec["error"] = f"rc={r.returncode}: {(r.stdout + r.stderr)[-300:]}"
                out.append(rec)
                continue
            lines = r.stdout.strip().splitlines()
            fwd, tot = (float(x) for x in lines[-1].split(","))
            chk = [l.split() for l in lines if l.startswith("CHECK 1 ")]
            rec.update(fwd_ms=round(fwd * 1e3, 3), fwd_train_ms=round(tot * 1e3, 3), startup_s=round(wall - 100 * tot, 2),
                       checksum=float(chk[0][2]) if chk else None, loss=float(chk[0][3]) if chk else None)
            out.append(rec)
    if not keep_data:
        subprocess.run(["rm", "-rf", data])
    return out


def main():
    ds, programs = sys.argv[1], sys.argv[2:]
    print(f"{'program':34s} {'gen':5s} {'fwd ms':>9s} {'fwd+train ms':>13s} {'start-up s':>11s}  epoch-1 checksum / loss")
    for rec in run(ds, programs, log=sys.stderr):
        if "error" in rec:
            print(f"{rec['program']:34s} {rec['generator']:5s} FAILED {rec['error']}")
        else:
            print(f"{rec['program']:34s} {rec['generator']:5s} {rec['fwd_ms']:9.3f} {rec['fwd_train_ms']:13.3f} "
                  f"{rec['startup_s']:11.2f}  {rec['checksum']:.6f} / {rec['loss']:.8f}", flush=True)


if __name__ == "__main__":
    main()
