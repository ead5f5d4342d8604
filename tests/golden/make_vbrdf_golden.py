"""Extract the GISS-Cox-Munk block (kernel index 10) of VLIDORT 2.8.3's saved BRDF self-test results into a small
fixture.  Run in the build container only (reads /root/reference):  python tests/golden/make_vbrdf_golden.py
Source: src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test/saved_results/gfortran/nstokes3/results_brdf_self_tester.res
Test setup (vlidort_v_test/V2p8p3_BRDF_self_tester.f90): nstokes 3, 1 quadrature stream, SZA 35, VZA 10,
azimuths 0/79/180, 100 azimuth-quadrature streams, shadowing on, factor 1,
pars = (0.5*(0.003+0.00512*15), refractive index of water at 0.5 um / 34.3 ppt)."""
import re
import numpy as np

SRC = ("/root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test/saved_results/gfortran/nstokes3/"
       "results_brdf_self_tester.res")


def main():
    txt = open(SRC).read()
    blk = txt.split("VBRDF Index: 10")[1].split("VBRDF Name")[0]
    nums = [float(x) for x in re.findall(r"-?\d\.\d{7}E[+-]\d\d", blk)]
    a = np.array(nums)
    assert a.size == 27 + 2 * 36, a.size
    exact = a[:27].reshape(3, 3, 3)                    # [azimuth][row o1][col o2]
    four = a[27:].reshape(2, 4, 3, 3)                  # [m][str-str, sun-str, str-user, sun-user][o1][o2]
    np.savez("tests/golden/vbrdf_self_tester_giss.npz", exact=exact, fourier=four)
    print("ok", exact[1], four[1, 3])


if __name__ == "__main__":
    main()


def supcheck_obsgeo():
    """3-kernel (Ross-thin 0.3, Li-dense 0.1 [2,1], Cox-Munk 0.1 [0.0798,1.779556] + shadow) scalar supplement check,
    observation geometry (35,10,10) (67,20,90) (75,40,170), 8 streams, 100 azimuth streams (half range -> 50)."""
    src = ("/root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test/saved_results/gfortran/obsgeo/"
           "results_brdf_supcheck_obsgeo.res")
    lines = [l for l in open(src).read().splitlines() if l.strip()]
    exact = np.array([float(l.split()[3]) for l in lines[3:6]])
    body = [l for l in lines[6:] if "Fourier" not in l]
    per = 8 * 2 + 3 + 1
    assert len(body) == 16 * per, len(body)
    F = np.zeros((16, 8, 8)); F0 = np.zeros((16, 8, 3)); UF = np.zeros((16, 3, 8)); UF0 = np.zeros((16, 3))
    for m in range(16):
        blk = body[m * per:(m + 1) * per]
        for j in range(8):
            F[m, j] = [float(x) for x in blk[2 * j].split()[1:]]
            F0[m, j] = [float(x) for x in blk[2 * j + 1].split()[1:]]
        for um in range(3):
            UF[m, um] = [float(x) for x in blk[16 + um].split()[1:]]
        UF0[m] = [float(x) for x in blk[19].split()[1:]]
    np.savez("tests/golden/vbrdf_supcheck_obsgeo.npz", exact=exact, brdf_f=F, brdf_f_0=F0, user_brdf_f=UF, user_brdf_f_0=UF0)
    print("supcheck ok", exact)


if __name__ == "__main__":
    supcheck_obsgeo()


def brdfplus():
    """TOA upwelling Stokes of VLIDORT's BRDF-plus tester (baseline run; solar-tester atmosphere, the 3-kernel
    BRDF above, regular-PS FO correction, delta-M, 8 streams): scalar obs-geometry and nstokes=3 lattice files."""
    base = "/root/reference/src/Components/rtms/RTSI/VLIDORT2p8p3/vlidort_v_test/saved_results/gfortran/"

    def toa(path, ngeom):
        out, cur = [], None
        for l in open(path):
            if "STOKES-" in l:
                cur = []; out.append(cur)
            elif "FLUX" in l:
                cur = None
            elif cur is not None and "Upwelling @  0.00" in l:
                cur.append(float(l.split()[4]))
        a = np.array(out)
        assert a.shape[1] == ngeom, a.shape
        return a
    obs = toa(base + "obsgeo/results_brdfplus_tester_obsgeo_I000.all", 3)
    lat = toa(base + "nstokes3/results_brdfplus_tester_lattice_IQU0.all", 36)
    np.savez("tests/golden/vbrdfplus_toa.npz", obsgeo_i=obs, lattice_iqu=lat)
    print("brdfplus ok", obs.shape, lat.shape)


if __name__ == "__main__":
    brdfplus()
