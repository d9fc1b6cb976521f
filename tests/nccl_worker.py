"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU, NCCL).

Every rank builds the SAME model (oracle.make_state_dict, seed 0) and the same synthetic batch, keeps the molecules of
its own contiguous shard (parallel.shard_bounds / take_shard), runs forward + backward of the additive loss
h.sum() + X.pow(2).sum() on the CUDA path and all-reduces the flat gradient buffer (parallel.FlatGradBuffer - the only
collective of the data path).  Rank 0 then runs the un-sharded batch on its own GPU and compares."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import gotennet_b200 as g
    from oracle import gotennet_oracle as orc

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    cfg = orc.OracleConfig(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True, scale_edge=False)
    sd = orc.expand_aliases(orc.make_state_dict(cfg, seed=0))

    def model():
        m = g.GotenNetWrapper(n_atom_basis=64, n_interactions=3, lmax=2, sep_dir=True, sep_tensor=True, scale_edge=False,
                              cutoff_fn=g.CosineCutoff(5.0), activation="swish")
        m.load_state_dict(sd, strict=True)
        return m.to(dev)

    class D:
        pass

    def step(m, z, pos, batch):
        d = D()
        d.z, d.pos, d.batch = z.to(dev), pos.to(dev), batch.to(dev)
        h, X = m(d)
        loss = h.sum() + X.pow(2).sum()
        loss.backward()
        return loss.detach()

    n_mol = 48
    z, pos, batch = orc.synth_batch("qm9", n_mol, seed=4)
    n_atoms = torch.bincount(batch, minlength=n_mol)
    lo, hi = g.shard_bounds((n_atoms.double() ** 2).tolist(), world)[rank]
    m = model()
    loss = step(m, *g.take_shard(z, pos, batch, lo, hi))
    fb = g.FlatGradBuffer(m.parameters())
    fb.all_reduce()
    dist.all_reduce(loss)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        ref = model()
        loss_ref = step(ref, z, pos, batch)
        fr = g.FlatGradBuffer(ref.parameters())
        fr.pack()
        err = (fb.flat - fr.flat).abs().max().item() / fr.flat.abs().max().item()
        worst = 0.0
        for a, b in zip(fb.views, fr.views):  # per tensor, at the parity bar
            worst = max(worst, (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30))
        lerr = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
        print(f"NCCL_GRAD_CHECK world={world} flat_rel={err:.3e} worst_tensor_rel={worst:.3e} loss_rel={lerr:.3e}", flush=True)
        ok = worst < 1e-4 and lerr < 1e-5
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
