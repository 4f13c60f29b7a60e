"""CLI of the hot path, flag-compatible with the reference's ``src/main.py`` (``main()``, 1073-1134).

    python -m ss25_hierarchical_multiscale_image_classification_b200.main --patch [--patch_level 0|1|2|3|all]
    python -m ss25_hierarchical_multiscale_image_classification_b200.main --extract_features [--patch_level L]
    python -m ss25_hierarchical_multiscale_image_classification_b200.main --patch --extract_features --fused

``--patch`` / ``-p``, ``--patch_level`` and ``--extract_features`` behave as in the reference (same
directory layout under ``./data/camelyon16``, same stage gating, same unknown-flag check).  ``--fused``
(new) skips the PNG round trip: slides go straight through the in-HBM pipeline and only the artefact files
are written.  The reference's other flags (download, training, evaluation...) are outside this package's
scope and are rejected by the unknown-flag check.
"""
from __future__ import annotations

import argparse
import os
import sys

from .feature_extraction import extract_features, extract_features_from_slides
from .preprocessing.extract import bcolors, extract_patches


def images_downloaded():
    img_dir = os.path.join(os.getcwd(), "data", "camelyon16", "train", "img")
    return os.path.exists(img_dir) and len([f for f in os.listdir(img_dir) if f.endswith(".tif")]) > 0


def patches_extracted(patch_level):
    patch_dir = os.path.join(os.getcwd(), "data", "camelyon16", "patches", f"level_{patch_level}")
    return os.path.exists(patch_dir) and any(os.listdir(patch_dir))


def features_extracted(patch_level):
    return os.path.exists(f"patch_features_{patch_level}.npy") and os.path.exists(f"patch_labels_{patch_level}.npy")


def main(argv=None, slide_opener=None):
    argv = sys.argv[1:] if argv is None else argv
    parser = argparse.ArgumentParser(description="Camelyon Dataset Processing (B200 hot path)")
    parser.add_argument("-p", "--patch", action="store_true", help="Extract patches")
    parser.add_argument("--patch_level", type=str, default="3",
                        help="WSI level for patch extraction (0, 1, 2, 3, or 'all' for all levels)")
    parser.add_argument("--extract_features", action="store_true", help="Extract features from patches")
    parser.add_argument("--fused", action="store_true", help="slides -> features in HBM, no PNG files")
    parser.add_argument("--device", type=str, default="cuda")
    known_args = {action.dest for action in parser._actions}
    input_args = {arg.lstrip('-').replace('-', '_') for arg in argv if arg.startswith('-')}
    unknown_args = input_args - known_args - {"p", "h"}
    if unknown_args:
        print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Unknown command line arguments: {', '.join(sorted(unknown_args))}")
        sys.exit(1)
    args = parser.parse_args(argv)
    levels = [0, 1, 2, 3] if args.patch_level == "all" else [int(args.patch_level)]

    if args.fused:
        if not images_downloaded():
            print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Images must be downloaded before extracting patches.")
            return
        for lvl in (levels if args.patch_level != "all" else [3]):
            extract_features_from_slides(level=lvl, slide_opener=slide_opener, device=args.device)
        return

    if args.patch:
        if not images_downloaded():
            print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Images must be downloaded before extracting patches.")
            return
        for lvl in levels:
            extract_patches(level=lvl, slide_opener=slide_opener, device=args.device)

    if args.extract_features:
        for lvl in levels:
            if not patches_extracted(lvl):
                print(f"{bcolors.ERROR}[ERROR]{bcolors.ENDC} Patches must be extracted at level {lvl} before extracting features.")
                return
        extract_features(level=int(args.patch_level) if args.patch_level != "all" else 3, device=args.device)


if __name__ == "__main__":
    main()
