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


def main():
    ds, programs = sys.argv[1], sys.argv[2:]
    data = os.path.join(CG, "_models", "Data", ds)
    subprocess.run(["rm", "-rf", data])
    subprocess.run([sys.executable, os.path.join(CG, "make_npy_dataset.py"), data, *SHAPES[ds]], check=True)
    print(f"{'program':34s} {'gen':5s} {'fwd ms':>9s} {'fwd+train ms':>13s} {'start-up s':>11s}  epoch-1 checksum / loss")
    for prog in programs:
        for kind in ("ref", "b200"):
            cwd = os.path.join(CG, "_models", f"{prog}_{kind}", "build")
            env = dict(os.environ)
            t0 = time.perf_counter()
            r = subprocess.run(["./gala_model"], cwd=cwd, capture_output=True, text=True, env=env)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                print(f"{prog:34s} {kind:5s} FAILED rc={r.returncode}: {(r.stdout + r.stderr)[-300:]}")
                continue
            lines = r.stdout.strip().splitlines()
            fwd, tot = (float(x) for x in lines[-1].split(","))
            chk = [l.split() for l in lines if l.startswith("CHECK 1 ")]
            c = f"{float(chk[0][2]):.6f} / {float(chk[0][3]):.8f}" if chk else "-"
            print(f"{prog:34s} {kind:5s} {fwd * 1e3:9.3f} {tot * 1e3:13.3f} {wall - 100 * tot:11.2f}  {c}", flush=True)
    subprocess.run(["rm", "-rf", data])


if __name__ == "__main__":
    main()
