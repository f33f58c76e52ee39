"""torchrun worker for tests/test_dist_nccl_gpu.py (and tools): N-rank prefill vs 1-rank prefill on the same box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/dist_worker.py --case mini|c3cut --out result.json

Every rank builds the same seeded weights and inputs, takes its shard (engine.make_plan) and runs ``Vidi15Engine.prefill`` at world N
-- once with the peer-memory exchange ("p2p") and once with the NCCL all-gather ("nccl").  Rank 0 then runs the SAME engine as a
single rank over the whole input and writes the comparison (replacement of Gather.forward, all_to_all.py:361).
Not a pytest module; nothing here reads /root/reference or calls the oracle."""
import argparse
import copy
import dataclasses
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="mini")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vidi_b200 import synth
    from vidi_b200.config import vidi15_mini, vidi15_true_dims
    from vidi_b200.engine import Vidi15Engine, make_plan
    if args.case == "mini":
        cfg = vidi15_mini()
        F, Cn, n_text = 2 * world + 1, max(2, world - 1), 13          # uneven frames; some ranks hold no audio chunk
        asz = Cn * 3000 - 1700                                         # partial last chunk
    else:
        # BASELINE config 3's shape at true 9B hidden dims with the depth cut: the 10x10 feature-map floor of the resize branch
        # (multimodal.py:175-180), >= 2 audio chunks per rank, uneven frame split
        cfg = dataclasses.replace(vidi15_true_dims(llm_layers=3, vis_layers=2, aud_layers=1, vocab=4096), max_image_tokens=300)
        F, Cn, n_text = 4 * world + 5, 2 * world + 1, 32
        asz = Cn * 3000 - 900
        assert cfg.image_hw(F) == (10, 10), cfg.image_hw(F)
    sd = synth.make_state_dict(cfg, seed=1234)
    sd = {k: (v if "mm_rand_pos" in k else v.to(BF).float()) for k, v in sd.items()}
    ids, images, mels, asz = synth.make_inputs(cfg, F, Cn, n_text=n_text, audio_size=asz)
    ids_dev = ids[ids != -200].cuda()
    images, mels = images.cuda().to(BF), mels.cuda().to(BF)
    plan = make_plan(cfg, F, Cn, asz, rank, world)
    res = dict(case=args.case, world=world, frames=F, chunks=Cn, audio_size=asz, image_hw=list(plan.hw))
    eng = None
    for mode in ("p2p", "nccl"):
        if eng is None:
            eng = Vidi15Engine(cfg, {k: v.clone() for k, v in sd.items()}, device=f"cuda:{local}", rank=rank, world=world,
                               group=None, exchange=mode)
        else:                      # same weights, other exchange
            eng = copy.copy(eng)
            eng.xchg, eng.exchange_note = None, "nccl all-gather of the per-rank reduced (O, LSE) blocks"
        logits = eng.prefill(ids_dev, images[plan.f0:plan.f1], mels[plan.c0:plan.c1], asz, n_frames_total=F, n_chunks_total=Cn)
        torch.cuda.synchronize()
        # every rank must hold the same logits: compare against rank 0's
        ref0 = logits.clone()
        dist.broadcast(ref0, 0)
        same = torch.tensor([1 if torch.equal(ref0, logits) else 0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        res[mode] = dict(exchange=eng.exchange_note, ranks_bit_equal=bool(same.item()), used=("p2p" if eng.xchg is not None else "nccl"))
        if rank == 0:
            one = copy.copy(eng)
            one.rank, one.world, one.xchg = 0, 1, None
            full = one.prefill(ids_dev, images, mels, asz)
            torch.cuda.synchronize()
            top2 = full.topk(2, -1).values
            err = float((logits - full).abs().max())
            decisive = (top2[:, 0] - top2[:, 1]) > 4 * err
            res[mode].update(rel_l2=rel(logits, full), max_abs=err, decisive_positions=int(decisive.sum()),
                             argmax_equal_on_decisive=bool(torch.equal(logits.argmax(-1)[decisive], full.argmax(-1)[decisive])),
                             argmax_equal_all=bool(torch.equal(logits.argmax(-1), full.argmax(-1))))
        dist.barrier()
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
