#!/usr/bin/env python
"""Drop-in for ``python retrieval/sample_retrieval.py`` (reference CLI at ``sample_retrieval.py:1673-1747``)
for ``Random`` and the ranked sampling methods: ``T2T-rank``, ``T2T-rank-T2I-tshd``, ``T2I-rank``, ``I2I-rank``,
``I2T-rank``, ``T2T-rank-I2T-tshd``, ``T2T-rank-I2I-tshd``, with ``--zeroshot_img_filter`` / ``--image_dedup``.

Same flags and defaults, same outputs: ``output/{dataset}_{model_cfg}_{prefix}/{prefix}.txt``
(``"<path> <label> 0"`` per accepted row, class-major), ``{prefix}_num_imgs_sampled.json``,
``sampling.log``, and a copy of the txt in ``../data/{dataset}/``.  Additive flags: ``--bank_dtype``
(``f32`` keeps the reference's fp32 features: the tcgen05 scan converts them to bf16 on the fly and every
candidate is re-scored exactly in fp32; ``bf16`` rounds the features once and halves the bytes streamed), ``--prompt_tensors`` (a cached prompt-tensor ``.pth`` as written by ``cal_prompt_tensors``
``:1433-1450``; the OpenCLIP text encoder that produces it is outside this package), ``--mined_pth``
/ ``--flat_shard`` to point at the feature file directly.
"""
import argparse
import json
import logging
import os
import random
import sys
from time import time

import torch

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swat_b200 import retrieval, shards  # noqa: E402
from swat_b200.config import retrieved_root  # noqa: E402


def get_logger(folder, name, mode="both"):
    """Same sinks as ``utils/logger.py:55-85``: ``{folder}/{name}.log`` (mode 'w') and/or console."""
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    logger.handlers.clear()
    fmt = logging.Formatter("%(asctime)s %(message)s")
    if mode in ("file", "both"):
        fh = logging.FileHandler(os.path.join(folder, f"{name}.log"), mode="w"); fh.setFormatter(fmt); logger.addHandler(fh)
    if mode in ("console", "both"):
        ch = logging.StreamHandler(); ch.setFormatter(fmt); logger.addHandler(ch)
    return logger


def build_parser():
    p = argparse.ArgumentParser(description="Arguments for script.")
    p.add_argument("--prefix", type=str, default=None, help="prefix.txt and foldername")
    p.add_argument("--dataset", type=str, default="semi-aves", help="Dataset name.")
    p.add_argument("--root", type=str, default=None, help="Root directory for storing mined data.")
    p.add_argument("--model_cfg", type=str, default="vitb32_openclip_laion400m",
                   choices=["vitb32_openclip_laion400m", "vitb32_openclip_laion2b", "vitb32_clip", "vitb16_clip"])
    p.add_argument("--database", type=str, default="LAION400M")
    p.add_argument("--prompt_name", type=str, default="alternates", choices=["most_common_name", "alternates", "name"])
    p.add_argument("--sampling_method", type=str, default="T2T-rank",
                   choices=["Random", "Random-I2I", "T2T-rank", "T2T-rank-T2I-tshd", "I2I-rank", "T2I-rank", "crossentropy",
                            "totalentropy", "I2T-rank", "I2T-tshd", "T2T-rank-I2T-tshd", "T2T-rank-I2I-tshd"])
    p.add_argument("--sampling_threshold", type=float, default=0.0)
    p.add_argument("--num_samples", type=int, default=500)
    p.add_argument("--zeroshot_img_filter", action="store_true", default=False)
    p.add_argument("--image_dedup", action="store_true", default=False)
    p.add_argument("--recal_prompt", action="store_true", default=False)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--log_mode", type=str, default="both", choices=["console", "file", "both"])
    # additive
    p.add_argument("--bank_dtype", type=str, default="f32", choices=["f32", "bf16"])
    p.add_argument("--prompt_tensors", type=str, default=None)
    p.add_argument("--mined_pth", type=str, default=None)
    p.add_argument("--flat_shard", type=str, default=None)
    p.add_argument("--caption_map_path", type=str, default=None)
    p.add_argument("--fewshot_path", type=str, default=None, help="few-shot feature .pth (default: the reference's ../data/{ds}/pre_extracted/...)")
    p.add_argument("--data_dir", type=str, default="../data", help="where the split txt is copied to (../data/{dataset}/)")
    p.add_argument("--device_index", type=int, default=0)
    return p


def main(argv=None):
    time_start = time()
    args = build_parser().parse_args(argv)
    if args.sampling_method not in ("Random", "T2T-rank", "T2T-rank-T2I-tshd", "T2I-rank", "I2I-rank", "I2T-rank",
                                    "T2T-rank-I2T-tshd", "T2T-rank-I2I-tshd"):
        raise NotImplementedError(f"--sampling_method {args.sampling_method} is outside the accelerated hot path; "
                                  "use the reference script for it")
    os.makedirs("output", exist_ok=True)
    random.seed(args.seed)
    torch.manual_seed(args.seed)
    args.case_name = f"{args.dataset}_{args.model_cfg}_{args.prefix}"
    args.output_folder = f"output/{args.case_name}"
    os.makedirs(args.output_folder, exist_ok=True)
    logger = get_logger(args.output_folder, "sampling", args.log_mode)
    logger.info(f"case_name: {args.case_name}")
    for arg in vars(args):
        logger.info(f"{arg} = {getattr(args, arg)}")
    root = args.root or retrieved_root()
    dataset_root = f"{root}/{args.dataset}"
    prompts_fn = args.prompt_tensors or os.path.join(args.data_dir, args.dataset, "prompts",
                                                     f"{args.dataset}_{args.model_cfg}_prompt_tensors.pth")
    if not os.path.exists(prompts_fn):
        raise FileNotFoundError(f"prompt tensors not found: {prompts_fn} (produce them with the reference's cal_prompt_tensors)")
    prompt_tensors_dict = torch.load(prompts_fn, map_location="cpu", weights_only=False)       # saved as CUDA tensors (:52)
    if args.prompt_name not in prompt_tensors_dict:              # a bare {cls: {'mean', 'all'}} dict
        prompt_tensors_dict = {args.prompt_name: prompt_tensors_dict}
    retrieval.prompt_tensors_dict = prompt_tensors_dict          # the reference's module global (:1740)
    feats = None
    if args.flat_shard:
        feats = shards.FlatShard(args.flat_shard).as_mined_dict()
    elif args.mined_pth:
        feats = shards.load_mined_pth(args.mined_pth)
    file_list_path, sample_ct = retrieval.sampling(args, logger, None, None, None, dataset_root, pre_extracted_feats=feats,
                                                   copy_to=os.path.join(args.data_dir, args.dataset))
    logger.info(f"sample_ct: {sample_ct}")
    logger.info(f"file_list_path: {file_list_path}")
    logger.info(f"Done, time: {round(time() - time_start)} seconds.")
    return file_list_path, sample_ct


if __name__ == "__main__":
    main()
