"""Output wire format of the reference's inference driver: what ``inf4eval.py`` writes for
``Evaluation/*.py`` to consume.

``save_batch_outputs`` mirrors ``DiFashion/inf4eval.py:774-827``: for every ``all_results[uid][oid]`` returned by
``fashion_generation(..., return_dict=False)`` it writes ``<gen_save_path>/images/<uid>/<oid>/<i>.jpg`` (one per generated
item, in generation order), ``all.jpg`` (GOR: the items pasted on a white ceil(sqrt(n))-column grid,
``merge_and_save_images`` ``:829-842``), replaces ``["images"]`` by ``["image_paths"]`` and accumulates the dictionary
that ``np.save(gen_save_path, np.array(outputs))`` stores (``:752``).  JPEG encoding is PIL's, as in the reference; the
pixels come from ``dfb_image_to_uint8``.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional, Sequence

import numpy as np


def _as_pil(img):
    from PIL import Image
    return img if isinstance(img, Image.Image) else Image.fromarray(np.asarray(img, dtype=np.uint8))


def merge_and_save_images(images: Sequence, save_path: str) -> None:
    """inf4eval.py:829-842."""
    from PIL import Image
    images = [_as_pil(i) for i in images]
    cols = math.ceil(math.sqrt(len(images)))
    width, height = images[0].width, images[0].height
    merged = Image.new("RGB", (width * cols, height * cols), color=(255, 255, 255))
    for i, im in enumerate(images):
        merged.paste(im, ((i % cols) * width, (i // cols) * height))
    merged.save(save_path)


def save_batch_outputs(all_outputs: Dict, all_grds: Dict, outputs: Dict, gen_save_path: str, task: str,
                       all_img_folder_path: Optional[str] = None, all_image_paths=None, test_grd_dict: Optional[Dict] = None,
                       save_grd: bool = True):
    """inf4eval.py:774-827.  ``outputs`` is consumed (its ``images`` entries are replaced by ``image_paths``)."""
    import torch
    for uid in outputs:
        for oid in outputs[uid]:
            imgs = [_as_pil(i) for i in outputs[uid][oid]["images"]]
            folder = os.path.join(gen_save_path, "images", str(uid), str(oid))
            os.makedirs(folder, exist_ok=True)
            if task == "GOR":
                merge_and_save_images(imgs, os.path.join(folder, "all.jpg"))
            paths = []
            for i, img in enumerate(imgs):
                path = os.path.join(folder, f"{i}.jpg")
                img.save(path)
                paths.append(path)
            outputs[uid][oid]["image_paths"] = paths
            del outputs[uid][oid]["images"]
            all_outputs.setdefault(uid, {}).setdefault(oid, outputs[uid][oid])
            if task == "FITB" and test_grd_dict is not None and all_img_folder_path is not None:
                from PIL import Image
                grd = [Image.open(os.path.join(all_img_folder_path, all_image_paths[iid])) for iid in test_grd_dict[oid]["outfits"]]
                merge_and_save_images(grd, os.path.join(folder, "grd.jpg"))
    if save_grd and test_grd_dict is not None:
        for uid in outputs:
            for oid in outputs[uid]:
                if oid in all_grds.get(uid, {}):
                    continue
                ent = all_grds.setdefault(uid, {}).setdefault(oid, {})
                ent["outfits"] = test_grd_dict[oid]["outfits"]
                paths = []
                for cate in outputs[uid][oid]["cates"]:
                    idx = torch.where(outputs[uid][oid]["full_cates"] == cate)[0]
                    iid = test_grd_dict[oid]["outfits"][idx]
                    paths.append(os.path.join(all_img_folder_path, all_image_paths[iid]))
                ent["image_paths"] = paths
    return all_outputs, all_grds


def save_outputs_npy(gen_save_path: str, all_outputs: Dict) -> str:
    """``np.save(gen_save_path, np.array(outputs))`` (inf4eval.py:752): a 0-d object array holding the dictionary,
    read back by the evaluation scripts with ``np.load(path, allow_pickle=True).item()``."""
    np.save(gen_save_path, np.array(all_outputs))
    return gen_save_path if gen_save_path.endswith(".npy") else gen_save_path + ".npy"
