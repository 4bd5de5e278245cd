"""Where the reference keeps the mined data and the caption maps
(``retrieval/extract_mined_feature.py:16-46``): paths under ``retrieved_path`` of ``config.yml`` whose
names encode how the images of a dataset were downloaded ("all" string-matched captions, or a random
subset of at most 2000 per class, ``laion_downloader.py:150``).  Resolved lazily so that importing
the package never needs ``config.yml``; ``SWAT_RETRIEVED_PATH`` overrides it."""
import os

# dataset -> download variant used by the reference for that dataset
DOWNLOAD_VARIANT = {"semi-aves": "all", "fgvc-aircraft": "all", "eurosat": "all", "dtd": "random", "flowers102": "random",
                    "oxford_pets": "random", "food101": "random", "stanford_cars": "random", "imagenet": "random"}


def retrieved_root(config_path: str = "../config.yml") -> str:
    root = os.environ.get("SWAT_RETRIEVED_PATH")
    if root:
        return root
    if os.path.exists(config_path):
        import yaml
        with open(config_path) as f:
            return yaml.safe_load(f)["retrieved_path"]
    return "."


class _PathTable(dict):
    def __init__(self, fmt):
        super().__init__()
        self.fmt = fmt

    def get(self, k, default=None):
        try:
            return self[k]
        except KeyError:
            return default

    def __missing__(self, k):
        if k not in DOWNLOAD_VARIANT:
            raise KeyError(k)
        return self.fmt.format(root=retrieved_root(), ds=k, var=DOWNLOAD_VARIANT[k])


MINED_DATASET_ROOT_DICT = _PathTable("{root}/{ds}/{ds}_retrieved_LAION400M-all_synonyms-{var}")     # :20-33
CAPTION_MAP_DICT = _PathTable("{root}/{ds}/{ds}_metadata-{var}-0.0-LAION400M.map")                  # :35-46
