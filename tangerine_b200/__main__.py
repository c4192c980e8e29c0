"""Headless export:  python -m tangerine_b200 export MODEL.tgm OUT.{ply,stl,vox} --grid CELLS_PER_UNIT [...]

The reference has no headless export flag (`--headless` only renders a frame); its export is
reachable through the dialog and through the legacy C entry points ExportPLY / ExportSTL / ExportMagicaVoxel
(tangerine/export.cpp:611-622, tangerine/magica.cpp:77-84).  This command is those three entry points on the CUDA
path (tg_export_ply / tg_export_stl / tg_export_magica_voxel), with the output format taken from the file extension
the way the dialog does it (tangerine/tangerine.cpp:895-918).  One deliberate difference: the reference's mesh export
ignores RefineIterations (only its point-cloud export refines, export.cpp:433-469) while tg_export_ply / tg_export_stl
honour it, so --refine defaults to 0 here (= the reference's files, byte for byte) and refinement is opt-in.  MODEL.tgm is a CSG tree dumped from the reference's Lua
front-end (oracle/ref_tool.cpp dump-tgm).  No CPU fallback: without a B200 the command fails with exit code 2.
"""
import argparse
import os
import sys

from . import api


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m tangerine_b200", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="command", required=True)
    ex = sub.add_parser("export", help="mesh (PLY / STL) or MagicaVoxel (.vox) export of a .tgm model")
    ex.add_argument("model")
    ex.add_argument("output")
    ex.add_argument("--grid", type=float, required=True, help="cells per model unit (GridSize of ExportPLY / ExportMagicaVoxel)")
    ex.add_argument("--refine", type=int, default=0,
                    help="refinement iterations applied to the MESH vertices (export.cpp:446-461).  Default 0 = byte-compatible with the "
                         "reference's ExportPLY / ExportSTL, whose MeshExportThread accepts RefineIterations and never uses it "
                         "(export.cpp:320-381); > 0 is this library's opt-in extension (BASELINE.json north_star)")
    ex.add_argument("--color-index", type=int, default=1, help="MagicaVoxel palette index (magica.cpp:62-66)")
    ex.add_argument("--device", type=int, default=0, help="CUDA device")
    info = sub.add_parser("info", help="bounds, primitive count and octree statistics of a .tgm model (host only)")
    info.add_argument("model")
    args = ap.parse_args(argv)
    try:
        tree = api.Tree.load(args.model)
        if args.command == "info":
            lo, hi = tree.bounds()
            print("bounds %s .. %s  primitives %d  painted %s" % (lo.tolist(), hi.tolist(), tree.leaf_count(), tree.has_paint()))
            print("octree %s" % tree.octree_stats())
            return 0
        ext = os.path.splitext(args.output)[1].lower()
        if ext == ".ply":
            api.export_ply(tree, args.grid, args.refine, args.output, args.device)
        elif ext == ".stl":
            api.export_stl(tree, args.grid, args.refine, args.output, args.device)
        elif ext == ".vox":
            api.export_magica_voxel(tree, args.grid, args.color_index, args.output, args.device)
        else:
            print("unknown export format %r: the reference writes .ply, .stl and .vox (tangerine/export.h:19-25)" % ext, file=sys.stderr)
            return 2
    except api.TangerineError as e:
        print(str(e), file=sys.stderr)
        return 2
    print("wrote %s (%d bytes)" % (args.output, os.path.getsize(args.output)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
