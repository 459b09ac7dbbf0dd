"""torchrun entry (one rank per GPU): partitioned GAT forward vs the single-GPU forward.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
          --master-port 29533 tests/dist_gpu_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gala-gnn-acceleration-language_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from gala_b200 import dist_gat, ops, synth  # noqa: E402
from gala_b200.gat_model import GAT2  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
n, e, feats = 50000, 3000000, 64
offset, ids = synth.powerlaw_csr_torch(n, e, seed=0, device=dev)
model = GAT2(feats, 32, 41, dev, seed=0)
X = torch.rand(n, feats, generator=torch.Generator(device=dev).manual_seed(1), device=dev) - 0.5
g = ops.TiledGraph(offset, ids, n).build_plan()
want = model.forward(g, X, mode="literal", dense="torch")
err = 0.0
for exchange in ("nccl", "p2p", "p2p-needed"):
    runner = dist_gat.PartitionedGAT(model, offset, ids, n, rank, world, dev, exchange=exchange)
    # "reflected": rows exchanged in the reflected basis (gala_gat_forward_col_f32; peer exchange only, NCCL -> folded)
    for mode in ("reflected", "folded"):
        for rep in range(3):    # repeated steps exercise the buffer re-use ordering of the peer exchange
            out_loc = runner.forward(X[runner.row_lo:runner.row_hi].contiguous(), mode=mode)
        full = runner.part.unpad(runner.part.all_gather(out_loc))
        e = float((full - want).double().norm() / want.double().norm())
        err = max(err, e)
        print(f"rank {rank}/{world} [{runner.exchange}, {mode}]: rows [{runner.row_lo},{runner.row_hi}) nnz "
              f"{runner.local_nvals} rel err {e:.3e}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if err < 1e-5 else 1)
