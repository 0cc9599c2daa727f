"""The reference's inference.py (29-154) on the sm_100a model: `test_inference`, `_sample_top_p`, `main`.

`test_inference` keeps the reference signature and loop semantics (prefill, then one token per step; greedy or
temperature + top-p; stop after EOS has been appended) but runs the loop on the device through
`PaliGemmaForConditionalGeneration.generate`.  `fire` is not installed in this image, so `main` is exposed through
argparse with the same flag names (`launch_inference.sh` flags work unchanged).
"""
import argparse

import torch

from . import _lib
from .modeling_gemma import KVCache  # noqa: F401  (re-exported like the reference module)
from .modeling_paligemma import PaliGemmaForConditionalGeneration


def move_inputs_to_device(model_inputs: dict, device: str):
    return {k: v.to(device) for k, v in model_inputs.items()}


def get_model_inputs(processor, prompt: str, image_file_path: str, device: str):
    from PIL import Image
    image = Image.open(image_file_path)
    model_inputs = processor(text=[prompt], images=[image])
    return move_inputs_to_device(model_inputs, device)


def _sample_top_p(probs: torch.Tensor, p: float, seed: int = 0, step: int = 0):
    """inference.py:90-106 on the device: probs [B, V] (CUDA) -> sampled token ids [B, 1] (int64).
    The kept set follows the reference rule exactly (exclusive cumulative mass <= p); the draw uses the kernel's
    counter-based RNG instead of torch.multinomial's CPU stream."""
    _lib.require_device()
    probs = probs.to(device="cuda", dtype=torch.float32)
    B, V = probs.shape
    logits = torch.log(probs).contiguous()  # softmax(log p) = p: the kernel normalises internally
    out = torch.empty(B, device="cuda", dtype=torch.int32)
    stp = torch.full((1,), int(step), device="cuda", dtype=torch.int32)
    _lib.check(_lib.lib().pg_sample_top_p(logits.data_ptr(), V, out.data_ptr(), 0, B, V, 1.0, float(p), int(seed),
                                          stp.data_ptr(), _lib.stream()), "pg_sample_top_p")
    return out.long().unsqueeze(-1)


def test_inference(model: PaliGemmaForConditionalGeneration, processor, device: str, prompt: str, image_file_path: str,
                   max_tokens_to_generate: int, temperature: float, top_p: float, do_sample: bool):
    model_inputs = get_model_inputs(processor, prompt, image_file_path, device)
    stop_token = processor.tokenizer.eos_token_id
    tokens = model.generate(model_inputs["input_ids"], model_inputs["pixel_values"], model_inputs["attention_mask"],
                            max_tokens_to_generate, do_sample=do_sample, temperature=temperature, top_p=top_p,
                            eos_token_id=stop_token)
    row = tokens[0]
    hits = (row == stop_token).nonzero()
    if hits.numel() > 0:  # the reference appends EOS, then breaks (inference.py:71-74)
        row = row[: int(hits[0]) + 1]
    decoded = processor.tokenizer.decode(row, skip_special_tokens=True)
    print(prompt + decoded)
    return row


def main(model_path: str = None, prompt: str = None, image_file_path: str = None, max_tokens_to_generate: int = 100,
         temperature: float = 0.8, top_p: float = 0.9, do_sample: bool = False, only_cpu: bool = False):
    if only_cpu:
        raise RuntimeError("this build has no CPU path: it needs a B200 (sm_100a)")
    from .processing_paligemma import PaliGemmaProcessor
    from .utils import load_hf_model
    device = "cuda"
    print("Device in use: ", device)
    print("Loading model")
    model, tokenizer = load_hf_model(model_path, device)
    model = model.to(device).eval()
    num_image_tokens = model.config.vision_config.num_image_tokens
    image_size = model.config.vision_config.image_size
    processor = PaliGemmaProcessor(tokenizer, num_image_tokens, image_size)
    print("Running inference")
    with torch.no_grad():
        test_inference(model, processor, device, prompt, image_file_path, max_tokens_to_generate, temperature, top_p, do_sample)


def _str2bool(v):
    return str(v).lower() in ("1", "true", "yes", "y")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_path")
    ap.add_argument("--prompt")
    ap.add_argument("--image_file_path")
    ap.add_argument("--max_tokens_to_generate", type=int, default=100)
    ap.add_argument("--temperature", type=float, default=0.8)
    ap.add_argument("--top_p", type=float, default=0.9)
    ap.add_argument("--do_sample", type=_str2bool, default=False)
    ap.add_argument("--only_cpu", type=_str2bool, default=False)
    main(**vars(ap.parse_args()))
