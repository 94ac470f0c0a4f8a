#!/bin/bash
# Builds the parts of the UNMODIFIED reference that compile without SuiteSparse (kvxopt.base, blas, lapack,
# misc_solvers + the pure-Python IPM) from the sources where they lie under /root/reference, against scipy's
# bundled OpenBLAS.  Output only under oracle/_ref/ (git-ignored, travels to the GPU box).  Used (a) to pin
# the oracle and the CUDA path against the reference's own LAPACK/IPM results, (b) as the harness in which the
# B200 cholmod/klu modules are plugged in as kvxopt.cholmod / kvxopt.klu (reference src/python/misc.py:21).
# The reference's cholmod.c / klu.c / umfpack.c need SuiteSparse headers and are NOT buildable here.
set -e
SRC=${REFERENCE_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref/kvxopt
[ -d "$SRC" ] || { echo "reference sources not present; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT"
OB=$(ls $(python3 -c "import scipy,os;print(os.path.dirname(scipy.__file__))")/../scipy.libs/libscipy_openblas-*.so | head -1)
OB=$(readlink -f "$OB")
REDEF=$HERE/_ref/redef.h
( grep -ho "\b[a-z][a-z0-9]*_\b *(" $SRC/C/{blas,lapack,base,dense,sparse,misc_solvers}.c | tr -d ' (' ;
  grep -o "define [a-z0-9]*_ " $SRC/C/blas_redefines.h | awk '{print $2}' ) | sort -u \
  | grep -E "^(d|z|i|dz|zd)[a-z0-9]+_$" | grep -v "^double_\|^int_\|^init_\|^index_\|^is_\|^do_\|^id_" \
  | awk '{printf "#define %s scipy_%s\n",$1,$1}' > "$REDEF"
EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
[ -f "$PYINC/Python.h" ] || PYINC=/usr/include/python3.12
CF="-O2 -fPIC -shared -w -I$PYINC -I$SRC/C -include $REDEF"
LD="$OB -Wl,-rpath,$(dirname $OB) -lm"
gcc $CF -DBASE_MODULE $SRC/C/base.c $SRC/C/dense.c $SRC/C/sparse.c -o $OUT/base$EXT $LD
for m in blas lapack misc_solvers; do gcc $CF $SRC/C/$m.c -o $OUT/$m$EXT $LD; done
cp $SRC/python/*.py $OUT/
printf 'version = "1.3.2.2"\nversion_tuple = (1,3,2,2)\n__version__ = version\n' > $OUT/_version.py
# the reference's own test-suite and doc examples, so that they can be executed VERBATIM against the B200 modules on the GPU
# box (where /root/reference does not exist): copied next to the build output only, never into the tracked tree
REFROOT=$(dirname "$SRC")
mkdir -p $HERE/_ref/tests $HERE/_ref/examples
cp $REFROOT/tests/*.py $REFROOT/tests/*.mtx $REFROOT/tests/*.mps $HERE/_ref/tests/
cp -r $REFROOT/examples/doc $HERE/_ref/examples/
echo "built reference probe in $OUT"
