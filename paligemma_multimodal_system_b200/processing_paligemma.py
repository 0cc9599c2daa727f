"""PaliGemmaProcessor with the reference's interface (processing_paligemma.py:94-212): upstream of the hot path
(SURVEY.md 8(f) "next" rows 3-4), kept as thin host code so that `inference.test_inference` is usable with a real
tokenizer.  Image path: bicubic resize -> /255 -> (x - 0.5) / 0.5 -> CHW float32; text path:
`<image>` * N + bos + prompt + "\\n" (processing_paligemma.py:77-89), tokenised by the caller's HF tokenizer.

Deliberate fixes relative to the reference, both noted in SURVEY.md: batches of prompts/images are accepted (the
reference asserts exactly one, :174), and each prompt string itself is formatted into the template (the reference
formats the *list*, :197, producing "...<bos>['caption en']\\n").
"""
from typing import List, Optional, Union

import numpy as np
import torch

IMAGENET_STANDARD_MEAN = [0.5, 0.5, 0.5]
IMAGENET_STANDARD_STD = [0.5, 0.5, 0.5]


def resize(image, resampling, image_size: int, reducing_gap: Optional[int] = None):
    return image.resize((image_size, image_size), resample=resampling, reducing_gap=reducing_gap)


def rescale(image: np.ndarray, scale_factor: float, dtype=np.float32) -> np.ndarray:
    return (image * scale_factor).astype(dtype)


def normalise(image: np.ndarray, mean: Union[float, List[float]], std: Union[float, List[float]]) -> np.ndarray:
    mean = np.array(mean, dtype=image.dtype)
    std = np.array(std, dtype=image.dtype)
    return (image - mean) / std


def process_images(images, image_size: int, scale_factor: float, resampling=None, reducing_gap: Optional[int] = None) -> List[np.ndarray]:
    out = []
    for image in images:
        image = resize(image=image, image_size=image_size, resampling=resampling, reducing_gap=reducing_gap)
        arr = np.array(image.convert("RGB"))
        arr = normalise(rescale(arr, scale_factor), IMAGENET_STANDARD_MEAN, IMAGENET_STANDARD_STD)
        out.append(arr.transpose(2, 0, 1))
    return out


def create_gemma_string(prefix_prompt: str, image_seq_len: int, image_token: str, bos_token: str) -> str:
    return f"{image_token * image_seq_len}{bos_token}{prefix_prompt}\n"


class PaliGemmaProcessor:
    IMAGE_TOKEN = "<image>"

    def __init__(self, tokenizer, num_image_tokens: int, image_size: int):
        self.tokenizer = tokenizer
        self.image_seq_len = num_image_tokens
        self.image_size = image_size
        self._add_new_tokens_to_tokenizer()
        self.tokenizer.add_eos_token = False
        self.tokenizer.add_bos_token = False

    def _add_new_tokens_to_tokenizer(self):
        self.tokenizer.add_special_tokens({"additional_special_tokens": [self.IMAGE_TOKEN]})
        extra = [f"<seg{i:03d}>" for i in range(128)] + [f"<loc{i:04d}>" for i in range(1024)]
        self.tokenizer.add_tokens(extra)
        self.tokenizer.image_token_id = self.tokenizer.convert_tokens_to_ids(self.IMAGE_TOKEN)

    def __call__(self, images, text: List[str], padding: str = "longest", truncation: bool = True, device=None) -> dict:
        """`device="cuda"`: the image path (bicubic resize, rescale, normalise, CHW) runs on the GPU
        (image_preprocess.process_images_gpu, bit-exact with the PIL / numpy path) and `pixel_values` stays on the device."""
        from PIL import Image
        if len(images) != len(text) or len(text) == 0:
            raise AssertionError(f"need one prompt per image, got {len(images)} images and {len(text)} prompts")
        if device is not None and str(device).startswith("cuda"):
            from .image_preprocess import process_images_gpu
            pixel_values = process_images_gpu(images, self.image_size, scale_factor=1 / 255.0)
        else:
            pixel_values = process_images(images, self.image_size, scale_factor=1 / 255.0, resampling=Image.Resampling.BICUBIC)
            pixel_values = torch.tensor(np.stack(pixel_values, axis=0))
        strings = [create_gemma_string(p, self.image_seq_len, self.IMAGE_TOKEN, self.tokenizer.bos_token) for p in text]
        tokens = self.tokenizer(strings, return_tensors="pt", truncation=truncation, padding=padding)
        return {"pixel_values": pixel_values, **tokens}
