"""torchrun entry: trials sharded over WORLD_SIZE GPUs must reproduce the single-GPU run of the whole batch,
both with the split path (phase A | NCCL all-reduce | phase B) and with the in-kernel exchange over NVLink peer
memory (vjf_run_sharded).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from vjf_b200.distributed import ShardedVJF, shard_bounds
from vjf_b200.model import VJF

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# CS_D > 480 exercises the wide-observation path (csrc/wide.cu) and its push all-reduce over peer memory
D, d, R, H, Bg, T = int(os.environ.get("CS_D", 60)), int(os.environ.get("CS_XD", 3)), int(os.environ.get("CS_R", 20)), [int(os.environ.get("CS_H", 16))], int(os.environ.get("CS_B", 200)), int(os.environ.get("CS_T", 12))
for lik in ("poisson", "gaussian"):
    torch.manual_seed(7)  # identical data on every rank
    y = (torch.poisson(torch.full((T, Bg, D), 0.8)) if lik == "poisson" else torch.randn(T, Bg, D)).cuda()
    eps = torch.randn(T, 2, Bg, d).cuda()
    lo, hi = shard_bounds(Bg, world, rank)
    ys, es = y[:, lo:hi].contiguous(), eps[:, :, lo:hi].contiguous()
    # reference: the same model on one GPU with the whole batch (fused persistent kernel), two epochs
    torch.manual_seed(11)
    ref = VJF.make_model(D, d, 0, R, H, lik, lr=1e-3, max_trials=Bg)
    rmu, rlv, rlosses = ref.run(y, eps=eps)
    rmu2, rlv2, rlosses2 = ref.run(y, eps=eps)
    for mode in ("split", "fused"):
        torch.manual_seed(11)  # identical parameters on every rank
        m = VJF.make_model(D, d, 0, R, H, lik, lr=1e-3, max_trials=Bg)
        sh = ShardedVJF(m)
        if mode == "fused":
            sh.connect()  # in-kernel all-reduce over NVLink peer memory
        mu, lv, losses = sh.run(ys, eps=es)
        mu2, lv2, losses2 = sh.run(ys, eps=es)  # second launch: the exchange epochs continue
        torch.cuda.synchronize()
        e_mu = max((mu - rmu[:, lo:hi]).abs().max().item(), (mu2 - rmu2[:, lo:hi]).abs().max().item())
        e_lv = max((lv - rlv[:, lo:hi]).abs().max().item(), (lv2 - rlv2[:, lo:hi]).abs().max().item())
        e_loss = ((losses2 - rlosses2).abs() / (1 + rlosses2.abs())).max().item()
        e_state = (m._flat - ref._flat).abs().max().item()
        flat = m._flat.clone()
        others = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(others, flat)
        lockstep = all(torch.equal(o, others[0]) for o in others)
        st = m.status()
        print(f"[rank {rank}] D={D} {lik}/{mode} kind={m._lib.vjf_last_launch_kind()} status={st}: |mu| {e_mu:.2e} |logvar| {e_lv:.2e} loss rel {e_loss:.2e} "
              f"state {e_state:.2e} replicas identical: {lockstep}", flush=True)
        assert e_mu < 5e-4 and e_lv < 5e-4 and e_loss < 5e-4 and e_state < 5e-3 and lockstep and st == 0
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_OK")
