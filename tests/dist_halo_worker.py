"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU, NCCL): a training step of a mesh
partitioned over the ranks, halo rows moved by DistExchange (all_to_all over NVLink) or by the library's own NCCL
transport (mgn_dp_* / mgn_halo_exchange), must reproduce the unpartitioned step computed on every rank's own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]


def _dp_check(pkg, orc, comm, devc, rank, world):
    """Data-parallel step through the library's own transport: every rank trains on its own window; mgn_backward_dp
    must leave the MEAN gradient on every rank (batch-P SGD) and the same Adam-updated parameters, and the
    online-normaliser merge must equal a serial pass over all windows."""
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(devc)
    fails = []
    for mode in (pkg.COMPUTE_FP32, pkg.COMPUTE_BF16):
        pos, cells, nt = orc.cylinder_flow_mesh(17, 11)
        s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
        N, E = pos.shape[0], s.shape[0]
        model, ps, _ = pkg.build_model(9, 2, 2, 3, 128, 2, device=devc, compute_mode=mode)
        rng = np.random.default_rng(100 + rank)
        nf, ef, tgt = (rng.normal(size=(N, 9)).astype(np.float32), rng.normal(size=(E, 3)).astype(np.float32),
                       rng.normal(size=(N, 2)).astype(np.float32))
        mask = dev(orc.node_mask(nt, [0, 5]))
        graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
        mgn = pkg.GraphNetwork(model, ps.clone(), None, None, None, None)
        (g,), _ = pkg.step_(mgn, graph, dev(tgt), mask)
        g_mean = g.clone()
        dist.all_reduce(g_mean)
        g_mean /= world
        opt = pkg.Adam(1e-3)
        st_ref, st_dp = opt.setup(ps), opt.setup(ps)
        ps_ref = ps.clone()
        opt.update(st_ref, ps_ref, g_mean)
        (g_dp,), _ = pkg.step_dp_(mgn, graph, dev(tgt), mask, opt=opt, opt_state=st_dp, comm=comm, n_buckets=5)
        torch.cuda.synchronize()
        e_g = float((g_dp - g_mean).norm() / g_mean.norm())
        e_p = float((mgn.ps - ps_ref).norm() / (ps_ref - ps).norm())
        print(f"[rank {rank} dp mode {mode}] mean-gradient err {e_g:.2e} updated-parameter err {e_p:.2e}", flush=True)
        if not (e_g < 1e-6 and e_p < 1e-3):
            fails.append(("dp", mode, e_g, e_p))
    # normaliser merge
    prev = torch.zeros(2 * 3 + 2, device=devc)
    norm = pkg.NormaliserOnline(3, devc)
    x = torch.from_numpy(np.random.default_rng(7 + rank).normal(size=(50, 3)).astype(np.float32)).to(devc)
    norm(x)
    comm.allreduce_normaliser_(norm.state, prev)
    allx = [torch.from_numpy(np.random.default_rng(7 + q).normal(size=(50, 3)).astype(np.float32)) for q in range(world)]
    serial = orc.NormaliserOnline(3)
    for a in allx:
        serial(a.numpy())
    want = np.concatenate([serial.acc_sum, serial.acc_sum_sq, [serial.acc_count, serial.num_acc]]).astype(np.float32)
    torch.cuda.synchronize()
    if not np.allclose(norm.state.cpu().numpy(), want, rtol=1e-5, atol=1e-5):
        fails.append(("norm", norm.state.cpu().numpy(), want))
    return fails


def main():
    import mgn_oracle as orc
    import mgn_pkg
    pkg = mgn_pkg.pkg
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    devc = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=devc)
    transport = sys.argv[1] if len(sys.argv) > 1 else "torch"
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(devc)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    fails = []
    for mode, tol, gtol in ((pkg.COMPUTE_FP32, 2e-5, 2e-4), (pkg.COMPUTE_BF16, 2e-2, 5e-2)):
        rng = np.random.default_rng(5)
        pos, cells, nt = orc.cylinder_flow_mesh(23, 14)
        s, r = orc.shift_to_one_based(*orc.triangles_to_edges(cells))
        N, E, mps = pos.shape[0], s.shape[0], 4
        model, ps, _ = pkg.build_model(9, 2, 2, mps, 128, 2, device=devc, compute_mode=mode)
        nf = rng.normal(size=(N, 9)).astype(np.float32)
        ef = rng.normal(size=(E, 3)).astype(np.float32)
        tgt = rng.normal(size=(N, 2)).astype(np.float32)
        mask = orc.node_mask(nt, [0, 5])
        graph = pkg.FeatureGraph(dev(nf), dev(ef), dev(s), dev(r))
        mgn = pkg.GraphNetwork(model, ps, None, None, None, None)
        (g_ref,), loss_ref = pkg.step_(mgn, graph, dev(tgt), dev(mask))
        out_ref = model.forward(graph, ps).cpu().numpy()
        g_ref, loss_ref = g_ref.cpu().numpy(), float(loss_ref.cpu())
        part = pkg.build_partition_rank(N, s, r, world, rank)
        glob = part.local_nodes_global()
        m2 = pkg.Model(9, 3, 2, mps, 128, 2, compute_mode=mode)
        pm = pkg.PartitionedModel(m2, part, dev(nf[glob]), dev(ef[part.edge_ids]), device=devc)
        own_mask = (mask[(mask - 1 >= part.lo) & (mask - 1 < part.hi)] - part.lo).astype(np.int32)
        if transport == "abi":
            comm = pkg.Communicator.from_torch_distributed(devc)
            ex = pkg.AbiExchange(part, world, comm, m2)
        else:
            ex = pkg.DistExchange(part, world, devc, m2)
        grads, losses, outs = pkg.run_partitioned_step([pm], ps, [dev(tgt[glob])], [dev(own_mask)], len(mask), ex,
                                                       pkg.masked_mse_partial)
        g, loss = grads[0], losses[0]
        if transport == "abi":
            comm.allreduce_sum_(g)
            comm.allreduce_sum_(loss)
        else:
            dist.all_reduce(g)
            dist.all_reduce(loss)
        torch.cuda.synchronize()
        own = outs[0].cpu().numpy()[:part.n_own]
        e_out = rel(own, out_ref[part.lo:part.hi])
        e_loss = abs(float(loss.cpu()) - loss_ref) / abs(loss_ref)
        e_g = rel(g.cpu().numpy(), g_ref)
        print(f"[rank {rank}/{world} {transport} mode {mode}] halo rows {len(part.halo_global)} out {e_out:.2e} "
              f"loss {e_loss:.2e} grad {e_g:.2e}", flush=True)
        if not (e_out < tol and e_loss < tol and e_g < gtol):
            fails.append((mode, e_out, e_loss, e_g))
    if transport == "abi":
        fails += _dp_check(pkg, orc, comm, devc, rank, world)
    bad = torch.tensor([len(fails)], device=devc)
    dist.all_reduce(bad)
    dist.destroy_process_group()
    sys.exit(1 if int(bad) else 0)


if __name__ == "__main__":
    main()
