"""64 camera-sized images -> 224 px pixel_values: PIL / numpy path of the reference vs the GPU kernels (H2D included)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from PIL import Image
from paligemma_multimodal_system_b200.image_preprocess import process_images_gpu
from paligemma_multimodal_system_b200.processing_paligemma import process_images
rng = np.random.default_rng(0)
for (H, W) in ((480, 640), (1080, 1920)):
    arrs = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(64)]
    pil = [Image.fromarray(a) for a in arrs]
    t0 = time.perf_counter()
    ref = np.stack(process_images(pil, 224, 1 / 255.0, resampling=Image.Resampling.BICUBIC))
    t_cpu = time.perf_counter() - t0
    pinned = [torch.from_numpy(a).pin_memory() for a in arrs]
    process_images_gpu(pinned, 224); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = process_images_gpu(pinned, 224); torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    same = np.array_equal(out.cpu().numpy(), ref)
    print(f"{H}x{W} x64 -> 224: reference PIL/numpy path {t_cpu * 1e3:.1f} ms, GPU path incl. H2D {t_gpu * 1e3:.2f} ms, bit-exact {same}")
