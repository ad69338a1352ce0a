"""BASELINE.json configs[4]: the decoder's share of a full synthesizer forward.

Mirrors SynthesizerSVC.forward (vits/model/synthesizers/synthesizer_svc.py:70-88) with pre-extracted content features
(PreloadHubertContentEncoder, content_encoder.py:76-126): enc_p -> emb_g -> enc_q -> flow -> rand_slice_segments -> dec,
on synthetic inputs of the configs/base.json shapes (feat ~ N(0,1) [B, 1280, Ty], pitch ids U{1..511}, spec ~ U(0,1)
[B, 1025, Ty], 4 s utterances = 375 frames, 32-frame decoder segments).  The non-decoder parts are plain PyTorch on
the GPU (oracle/synth_parts.py, pinned to the reference classes: the reference tree does not exist on the GPU box);
the decoder is timed twice: vcvits_b200.Generator (bf16 mode) and the PyTorch/cuDNN oracle decoder on the same GPU.

    python bench.py --workload full_synth        (or: python tools/full_synth.py)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(args=None):
    import torch
    from oracle import hifigan_oracle as O
    from oracle import synth_parts as S
    from vcvits_b200 import Generator

    assert torch.cuda.is_available(), "needs a CUDA device"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    steps = getattr(args, "steps", 20) if args is not None else 20
    warmup = getattr(args, "warmup", 5) if args is not None else 5
    B, Ty = 16, 375                     # batch 16 (configs/48k_base.json:17), 4 s utterances
    c = S.BASE_SYNTH
    torch.manual_seed(1234)
    parts = S.GeneratorParts(c).to(dev).eval()
    for p in parts.flow.parameters():   # the coupling layers' post convs are zero at init: give them work to do
        if float(p.detach().abs().sum()) == 0:
            p.data.normal_(0, 0.01)
    sd = O.seeded_state_dict(O.BASE_CFG, 1234)
    dec_b200 = Generator(**O.BASE_CFG, mode="bf16")
    dec_b200.load_state_dict(sd)
    dec_b200 = dec_b200.to(dev)
    dec_torch = O.build(O.BASE_CFG, sd).to(dev)
    dec_torch.gin_channels = O.BASE_CFG["gin_channels"]
    gen = torch.Generator(device="cpu").manual_seed(5)
    feats = torch.randn(B, c["hubert_channels"], Ty, generator=gen).to(dev)
    pitch = torch.randint(1, c["num_pitch"], (B, Ty), generator=gen).to(dev)
    spec = torch.rand(B, c["spec_channels"], Ty, generator=gen).to(dev)
    lens = torch.full((B,), Ty, dtype=torch.long, device=dev)
    sid = torch.randint(0, c["n_speakers"], (B,), generator=gen).to(dev)

    def run(dec, n):
        names, events = [], []

        def tick(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            names.append(name)
            events.append(e)

        acc = {}
        for it in range(n):
            names.clear(); events.clear()
            with torch.no_grad():
                S.generator_forward(parts, dec, feats, lens, pitch, spec, lens, sid, timer=tick)
            torch.cuda.synchronize()
            for k in range(len(events) - 1):
                acc[names[k]] = acc.get(names[k], 0.0) + events[k].elapsed_time(events[k + 1])
        return {k: v / n for k, v in acc.items()}

    out = {}
    for label, dec in (("vcvits_b200", dec_b200), ("pytorch_gpu", dec_torch)):
        run(dec, warmup)
        t = run(dec, steps)
        total = sum(t.values())
        out[label] = {"ms": {k: round(v, 4) for k, v in t.items()}, "total_ms": round(total, 4),
                      "decoder_share": round(t["dec"] / total, 4)}
    line = {"metric": "decoder_share_of_generator_forward", "value": out["vcvits_b200"]["decoder_share"], "unit": "fraction",
            "n_gpus": 1, "steps": steps, "warmup": warmup, "higher_is_better": False, "dtype": "bf16 decoder / fp32 rest",
            "data": "synthetic",
            "config": {"workload": "full_synth", "cfg": "base.json", "batch_per_gpu": B, "utterance_frames": Ty,
                       "decoder_segment_frames": c["segment_frames"],
                       "note": "forward only (synthesizer_svc.py:70-88); non-decoder parts = PyTorch eager on the GPU"},
            "with_vcvits_b200_decoder": out["vcvits_b200"], "with_pytorch_gpu_decoder": out["pytorch_gpu"],
            "decoder_speedup_vs_pytorch_gpu": round(out["pytorch_gpu"]["ms"]["dec"] / out["vcvits_b200"]["ms"]["dec"], 3)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
