"""Unit-to-waveform driver, drop-in for examples/speech_to_speech/generate_waveform_from_code.py:56-146 (same flags): reads
``{sample_id}|{units}`` lines (the reduced, normalized units), runs CodeHiFiGANVocoder on the GPU and writes
``{results_path}/{i}_pred.wav`` at 16 kHz.  `python -m diffnorm_b200.vocode_cli --in-code-file ... --vocoder g.pt
--vocoder-cfg config.json --results-path out --dur-prediction`."""
from __future__ import annotations

import argparse
import json
import os

import numpy as np
import torch

from .vocoder import CodeHiFiGANVocoder, load_code


def dump_result(args, sample_id, wav: torch.Tensor, suffix: str = ""):
    """generate_waveform_from_code.py:22-30 (soundfile is not a dependency here: scipy writes the same 16 kHz float wav)."""
    from scipy.io import wavfile
    wavfile.write(os.path.join(args.results_path, f"{sample_id}{suffix}_pred.wav"), 16000, wav.detach().cpu().numpy().astype(np.float32))


def main(args):
    if not torch.cuda.is_available() or args.cpu:
        raise SystemExit("diffnorm_b200's vocoder runs on CUDA only (there is no CPU path)")
    with open(args.vocoder_cfg) as f:
        vocoder_cfg = json.load(f)
    vocoder = CodeHiFiGANVocoder(args.vocoder, vocoder_cfg)
    data = load_code(args.in_code_file, args.reduce, filter_score=args.filter_score)
    os.makedirs(args.results_path, exist_ok=True)
    for i, d in enumerate(data):
        x = {"code": torch.LongTensor(d).view(1, -1)}
        wav = vocoder(x, args.dur_prediction)
        dump_result(args, i, wav)
        if args.limit is not None and i >= args.limit:
            return


def cli_main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--in-code-file", type=str, required=True, help="one output waveform per line")
    p.add_argument("--vocoder", type=str, required=True, help="path to the CodeHiFiGAN checkpoint")
    p.add_argument("--vocoder-cfg", type=str, required=True, help="path to the CodeHiFiGAN config")
    p.add_argument("--results-path", type=str, required=True)
    p.add_argument("--dur-prediction", action="store_true", help="enable duration prediction (for reduced/unique code sequences)")
    p.add_argument("--speaker-id", type=int, default=-1)
    p.add_argument("--cpu", action="store_true")
    p.add_argument("--reduce", action="store_true", help="remove consecutive duplicates of the unit sequence")
    p.add_argument("--filter-score", type=float, default=None)
    p.add_argument("--limit", type=int, default=None)
    main(p.parse_args(argv))


if __name__ == "__main__":
    cli_main()
