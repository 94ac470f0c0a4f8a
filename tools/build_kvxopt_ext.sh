#!/bin/bash
# Builds the reference's OWN extension modules kvxopt.cholmod and kvxopt.klu (src/C/cholmod.c, src/C/klu.c, read where
# they lie under /root/reference -- nothing is copied into the tracked tree) against include/suitesparse_shim/*.h and
# links them to libb200sparse.so in place of SuiteSparse (reference setup.py:292-298,389-394 link -lklu / -lcholmod).
# Output: oracle/_ref/kvxopt/{cholmod,klu}<EXT_SUFFIX> next to the probe build of the rest of kvxopt (git-ignored; they
# travel to the GPU box with the snapshot).  After this `import kvxopt.cholmod` / `kvxopt.klu` are compiled C modules
# whose numeric work runs on the B200, and the unmodified misc.kkt_chol2 picks them up at `from kvxopt import cholmod`.
#
# One line of cholmod.c is changed on the fly (in a temporary copy): numeric() samples Common.status BEFORE calling
# cholmod_l_factorize (cholmod.c:361-362), so in this revision a non-positive-definite matrix does not raise from
# numeric() although the documentation (cholmod.c:308-310, doc/source/spsolvers.rst:650) and misc.kkt_chol2's
# try/except (misc.py:1427-1433) say it does; the sampled line is moved below the call.  KVXOPT_EXT_UNPATCHED=1 keeps
# the file byte-for-byte.
set -e
SRC=${REFERENCE_SRC:-/root/reference/src}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=$HERE/oracle/_ref/kvxopt
[ -d "$SRC" ] || { echo "reference sources not present; keeping prebuilt extension modules"; exit 0; }
[ -f "$OUT/base.py" ] || [ -n "$(ls $OUT/base*.so 2>/dev/null)" ] || bash $HERE/oracle/build_ref.sh
[ -f "$HERE/kvxopt_b200/libb200sparse.so" ] || (cd $HERE && python3 -m kvxopt_b200.build)
EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
[ -f "$PYINC/Python.h" ] || PYINC=/usr/include/python3.12
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
if [ -n "$KVXOPT_EXT_UNPATCHED" ]; then
  cp $SRC/C/cholmod.c $TMP/cholmod.c
else
  # move `status = Common.status;` from before to after `CHOL(factorize) (Ac, Lc, &Common);`
  awk '/^    status = Common.status;$/ && !moved {held=1; next}
       {print}
       /CHOL\(factorize\) \(Ac, Lc, &Common\);/ && held && !moved {print "    status = Common.status;"; moved=1}' $SRC/C/cholmod.c > $TMP/cholmod.c
  grep -q "status = Common.status;" $TMP/cholmod.c
fi
# cholmod.c calls dcopy_/zcopy_ (diag, cholmod.c:897-898): same BLAS and symbol renaming as the probe build (oracle/build_ref.sh)
OB=$(ls $(python3 -c "import scipy,os;print(os.path.dirname(scipy.__file__))")/../scipy.libs/libscipy_openblas-*.so | head -1)
OB=$(readlink -f "$OB")
CF="-O2 -fPIC -shared -w -I$PYINC -I$HERE/include/suitesparse_shim -I$SRC/C -I$HERE/include -include $HERE/oracle/_ref/redef.h"
LD="-L$HERE/kvxopt_b200 -lb200sparse -Wl,-rpath,\$ORIGIN/../../../kvxopt_b200 $OB -Wl,-rpath,$(dirname $OB) -lm"
gcc $CF $TMP/cholmod.c $HERE/kvxopt_b200/csrc/suitesparse_shim.c -o $OUT/cholmod$EXT $LD
gcc $CF $SRC/C/klu.c $HERE/kvxopt_b200/csrc/suitesparse_shim.c -o $OUT/klu$EXT $LD
echo "built $OUT/cholmod$EXT and $OUT/klu$EXT against the SuiteSparse shim + libb200sparse.so"
