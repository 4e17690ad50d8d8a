"""Command-line front-ends with the reference's flags, defaults, file naming and page order.

    unmicstWrapper.py   unmicstWrapper.py:5-90   (1-based --channel/--classOrder/--GPU, --tool dispatch)
    UnMicst1-5.py       UnMicst1-5.py:713-876    unmicst-solo   (default model nucleiDAPI1-5)
    UnMicst2.py         UnMicst2.py:692-835      unmicst-duo    (default model nucleiDAPILAMIN)
    UnMicst.py          UnMicst.py:544-678       unmicst-legacy (default model nucleiDAPI)
    UnMicstCyto2.py     UnMicstCyto2.py:679-827  UnMicstCyto2   (CytoplasmIncell2)

The network runs once per image for all classes on the GPU; the host only reads the channel
page, hands it over with a PreMap, requantises (the reference's double uint8 quantisation,
UnMicst1-5.py:848-853) and writes the BigTIFF pages.  Additive options: ``--modelsDir`` /
$UNMICST_MODELS (where the ``models/<name>`` folders live), ``--gpus N`` / $UNMICST_GPUS (tile-row
bands over N GPUs of the box), ``--precision``.
"""
from __future__ import annotations

import argparse
import os
import queue
import sys
import threading
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import prepost, tiffio
from .engine import device_count
from .unet2d import UNet2D

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@dataclass
class ToolSpec:
    name: str
    default_model: str
    multi_channel_arg: bool       # --channel nargs='+' (solo, duo) vs a single int (legacy, Cyto2)
    stretch: bool                 # feed rescale_intensity(...) output (all but solo)
    proper_stem: bool             # solo strips '.ome.tif' properly; others split at the first dot
    one_based_names: bool         # '_Probabilities_<ch+1>' (Cyto2 writes <ch>)
    qc_dir: bool                  # preview under <out>/qc (Cyto2 writes next to the maps)
    two_inputs: bool = False      # duo: DNA + lamin stack


TOOLS = {
    "unmicst-solo": ToolSpec("unmicst-solo", "nucleiDAPI1-5", True, False, True, True, True),
    "unmicst-duo": ToolSpec("unmicst-duo", "nucleiDAPILAMIN", True, True, False, True, True, two_inputs=True),
    "unmicst-legacy": ToolSpec("unmicst-legacy", "nucleiDAPI", False, True, False, True, True),
    # the reference's default here is 'nucleiDAPI', which cannot be loaded by this graph
    "UnMicstCyto2": ToolSpec("UnMicstCyto2", "CytoplasmIncell2", False, True, False, False, False),
}


def tool_parser(spec: ToolSpec) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog=spec.name)
    p.add_argument("imagePath", help="path to the .tif file")
    p.add_argument("--model", help="type of model. For example, nuclei vs cytoplasm", default=spec.default_model)
    p.add_argument("--outputPath", help="output path of probability map")
    if spec.multi_channel_arg:
        p.add_argument("--channel", help="channel to perform inference on", nargs="+", default=[0])
    else:
        p.add_argument("--channel", help="channel to perform inference on", type=int, default=0)
    p.add_argument("--classOrder", help="background, contours, foreground", type=int, nargs="+", default=-1)
    p.add_argument("--mean", help="mean intensity of input image. Use -1 to use model", type=float, default=-1)
    p.add_argument("--std", help="mean standard deviation of input image. Use -1 to use model", type=float, default=-1)
    p.add_argument("--scalingFactor", help="factor by which to increase/decrease image size by", type=float, default=1)
    p.add_argument("--stackOutput", help="save probability maps as separate files", action="store_true")
    p.add_argument("--GPU", help="explicitly select GPU", type=int, default=-1)
    p.add_argument("--outlier", help="map percentile intensity to max when rescaling intensity values. Max intensity as default",
                   type=float, default=-1)
    p.add_argument("--verbose", help="display error messages for debugging", action="store_true")
    # additive
    p.add_argument("--modelsDir", help="folder holding the models/<name> directories", default=None)
    p.add_argument("--gpus", help="number of GPUs to shard tile rows over (0 = all)", type=int, default=None)
    p.add_argument("--precision", choices=["default", "auto", "fp32", "split3", "single"], default="default")
    return p


def resolve_model_dir(model: str, models_dir: Optional[str]) -> str:
    roots = [models_dir, os.environ.get("UNMICST_MODELS"), os.path.join(REPO, "models")]
    for r in roots:
        if r and os.path.isdir(os.path.join(r, model)):
            return os.path.join(r, model)
    raise FileNotFoundError(f"model folder '{model}' not found; pass --modelsDir or set UNMICST_MODELS "
                            f"(looked in {[r for r in roots if r]})")


def split_name(file_name: str, proper: bool):
    """(stem, type) — UnMicst1-5.py:783-792 vs `fileName.split('.', 1)` (UnMicst.py:603-605)."""
    parts = file_name.split(os.extsep)
    if proper:
        if len(parts) < 2:
            raise NotImplementedError("Input filename has no extension")
        if parts[-2] == "ome":
            return os.extsep.join(parts[:-2]), os.extsep.join(parts[-2:])
        return os.extsep.join(parts[:-1]), parts[-1]
    two = file_name.split(os.extsep, 1)
    if len(two) < 2:
        raise NotImplementedError("Input filename has no extension")
    return two[0], two[1]


def read_channel(path: str, file_type: str, channel: int) -> np.ndarray:
    if file_type in ("ome.tif", "ome.tiff", "btf", "tif", "tiff"):
        return tiffio.read_page(path, int(channel))
    if file_type == "czi":                    # czifile.CziFile(...).asarray()[0, 0, channel, 0, 0, :, :, 0]  (UnMicst1-5.py:797-800)
        from . import czi
        return czi.read_channel(path, int(channel))
    if file_type == "nd2":
        raise NotImplementedError(".nd2 needs the nd2reader package (Nikon's container is not published), which this build does "
                                  "not bundle; convert to (OME-)TIFF")
    raise NotImplementedError(f"Don't know how to read image with extension .{file_type}")


def pick_devices(gpu_flag: int, gpus: Optional[int]) -> List[int]:
    n_env = os.environ.get("UNMICST_GPUS")
    if gpus is None and n_env:
        gpus = 0 if n_env == "all" else int(n_env)
    n_dev = device_count()
    if n_dev == 0:
        raise RuntimeError("no CUDA device visible: unmicst_b200 has no CPU path")
    if gpus is not None:
        n = n_dev if gpus == 0 else min(gpus, n_dev)
        return list(range(n))
    if gpu_flag == -1:
        print("automatically choosing GPU")
        return [-1]
    return [gpu_flag]


def run_tool(tool: str, argv: Optional[List[str]] = None) -> int:
    spec = TOOLS[tool]
    args = tool_parser(spec).parse_args(argv)
    model_dir = resolve_model_dir(args.model, args.modelsDir)
    devices = pick_devices(args.GPU, args.gpus)
    print("Using GPU " + (str(devices[0]) if len(devices) == 1 else str(devices)))
    precision = args.precision          # "auto" calibrates against the split mode on probe tiles (engine.Engine)
    UNet2D.singleImageInferenceSetup(model_dir, devices if len(devices) > 1 else devices[0], args.mean, args.std,
                                     precision=precision)
    try:
        return _process(spec, args)
    finally:
        UNet2D.singleImageInferenceCleanup()


def _process(spec: ToolSpec, args) -> int:
    n_class = UNet2D.hp["nClasses"]
    image_path = args.imagePath
    if spec.multi_channel_arg:
        chans = [int(c) for c in args.channel]
    else:
        chans = [int(args.channel)]
    dapi = chans[0]
    if spec.two_inputs:
        chans = [dapi, dapi] if len(chans) == 1 else chans[:2]
        print("Using channels " + str(chans[0] + 1) + " and " + str(chans[1] + 1))
    else:
        chans = chans[:1]
        print("Using channel " + str(dapi + 1))
    parent = os.path.dirname(os.path.dirname(image_path))
    stem, ftype = split_name(os.path.basename(image_path), spec.proper_stem)

    raws = [prepost.coerce_raw(read_channel(image_path, ftype, c)) for c in chans]
    raw_shape = raws[0].shape
    raw_last = raws[-1]                                       # rawI is the last page read (UnMicst2.py:771,792)
    # the raw samples go to the GPU untouched; img_as_float, resize (--scalingFactor) and the rescale_intensity stretch
    # (each channel with its own min/max) happen in the gather kernel, the resize back and the second uint8
    # quantisation in the output kernels (UnMicst1-5.py:813-821,848-853)
    prepared = [prepost.network_input(r, args.scalingFactor, spec.stretch, args.outlier, engine=UNet2D.Engine) for r in raws]
    infer_shape = prepared[0][2]
    if len(prepared) == 1:
        image, premap = prepared[0][0], prepared[0][1]
    else:
        if any(p[0].dtype != prepared[0][0].dtype or p[0].shape != raw_shape for p in prepared):
            raise ValueError("the channel pages differ in size or sample type")
        image, premap = np.stack([p[0] for p in prepared]), [p[1] for p in prepared]

    class_order = list(range(n_class)) if args.classOrder == -1 else list(args.classOrder)
    out_dir = args.outputPath if args.outputPath else parent + "//probability_maps"
    os.makedirs(out_dir, exist_ok=True)
    qc_dir = os.path.join(out_dir, "qc") if spec.qc_dir else out_dir
    if spec.qc_dir:
        os.makedirs(qc_dir, exist_ok=True)
    suffix = str(dapi + 1) if spec.one_based_names else str(dapi)
    preview = prepost.preview_page(raw_last)
    H, W = raw_shape

    # files and the class each of their pages holds; pages fill band by band while the GPU works on the next band
    # (a writer thread drains a queue), then the preview pages are appended
    if args.stackOutput:
        prob_path = os.path.join(out_dir, f"{stem}_Probabilities_{suffix}.tif")
        plan = [(prob_path, list(class_order[::-1]))]         # backwards in order to align with ilastik
        if len(class_order) > 1:
            plan.append((os.path.join(qc_dir, f"{stem}_Preview_{suffix}.tif"), [class_order[::-1][1]]))
    else:
        if len(class_order) < 3:
            raise ValueError("without --stackOutput the contours/nuclei pair needs a 3-class model "
                             "(the reference indexes classOrder[2] and fails too); use --stackOutput")
        plan = [(os.path.join(out_dir, f"{stem}_ContoursPM_{suffix}.tif"), [class_order[1]]),
                (os.path.join(out_dir, f"{stem}_NucleiPM_{suffix}.tif"), [class_order[2]])]
    writers = []
    for path, classes in plan:
        w = tiffio.BigTiffWriter(path, append=False)
        w.begin_pages(len(classes), H, W, np.uint8)
        writers.append((w, classes))
    bands: "queue.Queue" = queue.Queue(maxsize=4)
    errors: List[BaseException] = []

    def drain():
        while True:
            item = bands.get()
            if item is None:
                return
            r0, _, buf = item
            try:
                for w, classes in writers:
                    for i, cls in enumerate(classes):
                        w.write_page_rows(i, r0, buf[cls])
            except BaseException as ex:          # keep draining so the producer never blocks
                errors.append(ex)

    t = threading.Thread(target=drain)
    t.start()
    try:
        for band in UNet2D.singleImageInferenceStream(image, premap=premap, infer_shape=infer_shape, cli_quant=True):
            bands.put(band)
    finally:
        bands.put(None)
        t.join()
    if errors:
        raise errors[0]
    for w, _ in writers:
        w.end_pages()
    # the raw preview follows the contour page: qc/<stem>_Preview (stack mode) or <stem>_ContoursPM (UnMicst1-5.py:857-872)
    if len(writers) > 1 or not args.stackOutput:
        (writers[1][0] if args.stackOutput else writers[0][0]).write_page(preview)
    for w, _ in writers:
        w.close()
    return 0


# ------------------------------------------------------------------------------------------------
# unmicstWrapper.py
# ------------------------------------------------------------------------------------------------
def wrapper_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="unmicstWrapper.py")
    p.add_argument("--tool", help="which UnMicst tool?", default="unmicst-solo")
    p.add_argument("imagePath", help="path to the .tif file")
    p.add_argument("--model", help="type of model. For example, nuclei vs cytoplasm")
    p.add_argument("--outputPath", help="output path of probability map")
    p.add_argument("--channel", help="channel to perform inference on", nargs="+", type=int, default=[1])
    p.add_argument("--classOrder", help="background, contours, foreground", type=int, nargs="+", default=-1)
    p.add_argument("--mean", help="mean intensity of input image. Use -1 to use model", type=float, default=-1)
    p.add_argument("--std", help="mean standard deviation of input image. Use -1 to use model", type=float, default=-1)
    p.add_argument("--scalingFactor", help="factor by which to increase/decrease image size by", type=float, default=1)
    p.add_argument("--stackOutput", help="save probability maps as separate files", action="store_true")
    p.add_argument("--GPU", help="explicitly select GPU", type=int, default=0)
    p.add_argument("--outlier", help="map percentile intensity to max when rescaling intensity values. Max intensity as default",
                   type=float, default=-1)
    p.add_argument("--verbose", help="display error messages for debugging", action="store_true")
    p.add_argument("--modelsDir", default=None)
    p.add_argument("--gpus", type=int, default=None)
    p.add_argument("--precision", choices=["default", "auto", "fp32", "split3", "single"], default="default")
    return p


def wrapper_to_tool_argv(args) -> (str, List[str]):
    """The 1-based -> 0-based conversion and command assembly of unmicstWrapper.py:28-85."""
    channel = [c - 1 for c in args.channel]
    class_order = -1 if args.classOrder == -1 else [c - 1 for c in args.classOrder]
    gpu = args.GPU - 1
    if args.tool == "unmicst-duo":
        tool = "unmicst-duo"
        ch = [str(channel[0]), str(channel[1])] if len(channel) == 2 else [str(channel[0])]
    elif args.tool == "unmicst-legacy":
        tool = "unmicst-legacy"
        ch = [str(channel[0])]
        print("")
        print("WARNING! YOU HAVE OPTED TO USE UNMICST legacy, WHICH IS GETTING TIRED AND OLD. CONSIDER USING unmicst-solo "
              "OR unmicst-duo (IF YOU ALSO HAVE A NUCLEAR ENVELOPE STAIN")
        print("")
    elif args.tool == "UnMicstCyto2":
        tool = "UnMicstCyto2"
        ch = [str(channel[0])]
    else:
        tool = "unmicst-solo"
        ch = [str(channel[0])]
        print("")
        print("WARNING! USING unmicst-solo AS DEFAULT. THIS MODEL HAS BEEN TRAINED ON MORE TISSUE TYPES. IF YOU WANT THE "
              "LEGACY MODEL, USE --tool unmicst-legacy")
        print("")
    argv = [args.imagePath, "--channel"] + ch
    if args.outputPath is not None:          # the reference passes the literal string 'None' here
        argv += ["--outputPath", str(args.outputPath)]
    argv += ["--mean", str(args.mean), "--std", str(args.std), "--scalingFactor", str(args.scalingFactor),
             "--GPU", str(gpu), "--outlier", str(args.outlier)]
    if args.stackOutput:
        argv.append("--stackOutput")
    if args.model:
        argv += ["--model", str(args.model)]
    if class_order != -1:
        argv += ["--classOrder"] + [str(c) for c in class_order[:3]]
    if args.verbose:
        argv.append("--verbose")
    if args.modelsDir:
        argv += ["--modelsDir", args.modelsDir]
    if args.gpus is not None:
        argv += ["--gpus", str(args.gpus)]
    if args.precision != "default":
        argv += ["--precision", args.precision]
    return tool, argv


def run_wrapper(argv: Optional[List[str]] = None) -> int:
    args = wrapper_parser().parse_args(argv)
    tool, tool_argv = wrapper_to_tool_argv(args)
    print(" ".join([tool] + tool_argv))
    return run_tool(tool, tool_argv)     # in-process instead of os.execvp (unmicstWrapper.py:88-90 FIXME)


if __name__ == "__main__":
    sys.exit(run_wrapper())
