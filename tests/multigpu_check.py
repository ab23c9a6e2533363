"""Multi-GPU check of the data-parallel launcher over NCCL (run under torchrun on >= 2 GPUs; not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py

Every rank encodes its LPT share of a ragged batch, the final gather (NCCL all-gather-v + row gather, and the NVLink
peer-memory scatter kernel) restores the original order, and the gathered embeddings are checked bit-for-bit against each
other on every rank and against a single-GPU encode of the whole batch on rank 0.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from qwen3_asr_mlx_b200 import AudioEncoder, AudioEncoderConfig, launcher, weights  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = AudioEncoderConfig(encoder_layers=2)
    enc = AudioEncoder(cfg, device=local)
    enc.load_weights(weights.random_init(cfg, seed=1234))
    rng = np.random.default_rng(99)
    lengths = [int(n) for n in rng.integers(8000, 300000, size=23)]
    audios = [synth(np.random.default_rng(1000 + i), n) for i, n in enumerate(lengths)]

    def encode_fn(idx):
        emb, toffs = enc.encode_audio_batch([audios[i] for i in idx])
        return emb.tensor, toffs

    emb, offs, mine = launcher.encode_sharded(encode_fn, lengths, cfg.output_dim, rank, world, tokens_per_call=2048)
    # the same gather through NVLink peer memory (one scatter kernel, no NCCL collective on the data path), twice (buffer reuse)
    pg = launcher.PeerGather(int(offs[-1]) + 64, cfg.output_dim, dtype=torch.float32)
    for _ in range(2):
        emb_p2p, offs2, _ = launcher.encode_sharded(encode_fn, lengths, cfg.output_dim, rank, world, tokens_per_call=2048, peer_gather=pg)
    ok = bool(torch.equal(emb_p2p, emb)) and list(offs2) == list(offs)
    if rank == 0:
        ref, ref_offs = enc.encode_audio_batch(audios)
        ok = ok and bool(torch.equal(emb, ref.tensor)) and list(offs) == list(ref_offs)
        print(f"world={world} shares={[len(p) for p in launcher.lpt_partition([launcher.tokens_for_samples(n) for n in lengths], world)]} "
              f"tokens={int(offs[-1])} gathered (NCCL and peer-memory scatter) == single-GPU: {ok}")
    # contiguous token-balanced shares + overlapped NVLink DMA block gather (the path bench.py measures at N > 1), twice
    costs = [launcher.tokens_for_samples(n) for n in lengths]
    share = launcher.contiguous_partition(costs, world)[rank]
    packed = torch.from_numpy(np.concatenate([audios[i] for i in share])).cuda()
    for dt, name in ((torch.float32, "float32"), (torch.bfloat16, "bfloat16")):
        bg = launcher.PeerBlockGather(int(offs[-1]) + 16, cfg.output_dim, dtype=dt)
        for _ in range(2):
            emb_c, offs3, mine3 = launcher.encode_contiguous_sharded(enc, packed, lengths, rank, world, gather=bg, tokens_per_call=2048, out_dtype=name,
                                                                         tail_min_rows=256)  # exercises the tail-block projector path
        want = emb if dt == torch.float32 else None
        if want is None:  # bf16 output: compare with a bf16 single-GPU encode of the whole batch (every rank computes it)
            want = enc.encode_audio_batch(audios, out_dtype="bfloat16")[0].tensor
        ok_c = bool(torch.equal(emb_c, want)) and list(offs3) == list(offs) and mine3 == share
        if rank == 0:
            print(f"contiguous shares {[len(p) for p in launcher.contiguous_partition(costs, world)]} + DMA block gather ({name}) == single-GPU: {ok_c}")
        ok = ok and ok_c
        del bg
    # config 4, single pass: ONE long utterance, its attention windows sharded over the ranks (launcher.encode_long_sharded)
    from qwen3_asr_mlx_b200 import log_mel_spectrogram

    long_audio = synth(np.random.default_rng(4), 16000 * 95 + 777)
    mel = log_mel_spectrogram(long_audio).tensor          # every rank computes the full mel (utterance-wide max)
    whole = enc(mel).tensor[0]
    pg2 = launcher.PeerGather(int(whole.shape[0]) + 8, cfg.output_dim, dtype=torch.float32)
    sharded_nccl = launcher.encode_long_sharded(enc, mel, rank, world)
    sharded_p2p = launcher.encode_long_sharded(enc, mel, rank, world, peer_gather=pg2)
    bg2 = launcher.PeerBlockGather(int(whole.shape[0]) + 8, cfg.output_dim, dtype=torch.float32)
    sharded_dma = launcher.encode_long_sharded(enc, mel, rank, world, peer_gather=bg2)
    ok_long = bool(torch.equal(sharded_nccl, whole)) and bool(torch.equal(sharded_p2p, whole)) and bool(torch.equal(sharded_dma, whole))
    if rank == 0:
        print(f"config 4 single pass: {int(whole.shape[0])} tokens, window shares {launcher.window_shares(int(mel.shape[1]), world)}, "
              f"sharded == single-GPU: {ok_long}")
    ok = ok and ok_long
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
