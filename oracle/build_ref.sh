#!/bin/bash
# Recipe for oracle/_ref/mulskips.e: the REFERENCE ITSELF (Fortran, /root/reference/mulskips-source) built with gfortran, for
# boxes that have both a Fortran compiler and the reference tree.  TEST INFRASTRUCTURE, like everything under oracle/.
#
# Neither exists where this repo was built (no gfortran / f95 / flang in the image: probed; the GPU box has no /root/reference),
# so this script has never run: it records the recipe SURVEY.md 8(d) asks for and exits with a message when a piece is
# missing.  The sources are compiled from a scratch copy OUTSIDE the repo (the box size is a compile-time PARAMETER that has
# to be rewritten, py/setuprun.py:499-522); only the binary lands in oracle/_ref/ (git-ignored).
#
#   oracle/build_ref.sh [LenX LenY LenZ] [NumMcMax]      default 480 480 480, the reference's 500000
set -euo pipefail
REF=${MSKC_REFERENCE:-/root/reference/mulskips-source}
FC=${FC:-gfortran}
HERE=$(cd "$(dirname "$0")" && pwd)
LX=${1:-480}; LY=${2:-480}; LZ=${3:-480}; NMC=${4:-500000}
command -v "$FC" >/dev/null 2>&1 || { echo "build_ref: no Fortran compiler ($FC) on this box -- oracle/_ref cannot be built; the C restatement (oracle/mskc_oracle.c) is the CPU arm"; exit 3; }
[ -f "$REF/main.f" ] || { echo "build_ref: $REF/main.f not found -- the reference tree is not on this box"; exit 3; }
TMP=$(mktemp -d /tmp/mskc_ref.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF"/. "$TMP"/
# the box and the active-list cap are PARAMETERs of modules/defsystem.f (defsystem.f:31-34)
sed -i -E "s/(INTEGER, PARAMETER :: LenX=)[0-9]+(, *LenY=)[0-9]+(, *LenZ=)[0-9]+/\1${LX}\2${LY}\3${LZ}/" "$TMP/modules/defsystem.f"
sed -i -E "s/(NumMcMax *= *)[0-9]+/\1${NMC}/" "$TMP/modules/defsystem.f"
# module files first, then everything else (the order of the reference's own makefile:15-34)
SRC="modules/defdertype.f modules/defsystem.f modules/definitions.f $(cd "$TMP" && ls initial/*.f inpout/*.f lattice/*.f mcevent/*.f | tr '\n' ' ') main.f"
mkdir -p "$HERE/_ref"
( cd "$TMP" && "$FC" -O3 -march=native -mcmodel=medium -o "$HERE/_ref/mulskips.e" $SRC )
echo "build_ref: $HERE/_ref/mulskips.e built for a ${LX}x${LY}x${LZ} box (NumMcMax ${NMC})"
