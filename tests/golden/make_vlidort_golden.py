"""Generates tests/golden/vlidort_solar_tester.npz from the reference's own test inputs and saved results
(run in the build container, where /root/reference exists; the fixture is committed so that tests never
read /root/reference).

Sources (relative to /root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test/):
  input_atmos.dat, ProblemIII.Moms                  -- the 23-layer atmosphere of V2p8p3_solar_tester.f90:252-300
  saved_results/gfortran/results_solar_tester_I000.all            (nstokes=1)
  saved_results/gfortran/nstokes3/results_solar_tester_IQU0.all   (nstokes=3)
"""
import os
import re
import numpy as np

REF = "/root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vlidort_solar_tester.npz")


def fortran_floats(line, start, width, n):
    return [float(line[start + i * width:start + (i + 1) * width]) for i in range(n)]


def read_atmos(nlayers=23):
    with open(os.path.join(REF, "input_atmos.dat")) as f:
        lines = f.read().split("\n")
    raymoms = np.array([fortran_floats(lines[i], 5, 18, nlayers) for i in range(3)])     # (i5,1p25e18.9)
    height = np.zeros(nlayers + 1); height[0] = 60.0
    molext = np.zeros(nlayers); molomg = np.zeros(nlayers)
    for n in range(nlayers):
        ln = lines[3 + n]                                                               # (i4,f12.5,1p6e16.7)
        height[n + 1] = float(ln[4:16]); molext[n] = float(ln[16:32]); molomg[n] = float(ln[32:48])
    return raymoms, height, molext, molomg


def read_moms():
    with open(os.path.join(REF, "ProblemIII.Moms")) as f:
        lines = f.read().split("\n")[5:]
    # V2p8p3_solar_tester.f90:272-295 (cutoff logic)
    aermoms = []
    lcheck, cutoff = 0, 1.0e-06
    for ln in lines:
        if len(aermoms) - 1 >= 106 or lcheck >= 2:
            break
        t = ln.split()
        if len(t) < 7:
            break
        am = [float(x) for x in t[1:7]]
        if am[0] > cutoff:
            aermoms.append(am)
            if lcheck == 1:
                lcheck += 1
        else:
            lcheck += 1
            if lcheck == 1:
                aermoms.append(am)
    return np.array(aermoms).T      # (6, nmom+1)


def read_results(path, nblocks):
    """returns array (nblocks, 36 geometries, 6 tasks) of 'Upwelling @  0.00' (TOA) values"""
    vals = [[] for _ in range(nblocks)]
    blk = -1
    with open(path) as f:
        for ln in f:
            if re.search(r"STOKES-[IQUV] , tasks", ln):
                blk += 1
            m = re.match(r"\s*(\d+)\s+Upwelling @\s+0\.00\s+(.*)$", ln)
            if m and 0 <= blk < nblocks and len(vals[blk]) < 36:
                vals[blk].append([float(x) for x in m.group(2).split()])
    return np.array([np.array(v)[:36] for v in vals])


if __name__ == "__main__":
    raymoms, height, molext, molomg = read_atmos()
    aermoms = read_moms()
    gold1 = read_results(os.path.join(REF, "saved_results/gfortran/results_solar_tester_I000.all"), 1)
    gold3 = read_results(os.path.join(REF, "saved_results/gfortran/nstokes3/results_solar_tester_IQU0.all"), 3)
    np.savez(OUT, raymoms=raymoms, height=height, molext=molext, molomg=molomg, aermoms=aermoms,
             szas=np.array([35.0, 67.0, 75.0, 82.0]), vzas=np.array([10.0, 20.0, 40.0]),
             azms=np.array([10.0, 90.0, 170.0]), gold_i_ns1=gold1[0], gold_iqu_ns3=gold3)
    print(OUT, raymoms.shape, aermoms.shape, gold1.shape, gold3.shape)
