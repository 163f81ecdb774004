"""ctypes access to the internals of oracle/_ref/libklt_ref.so (the unmodified
reference CPU sources compiled in place).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C

import numpy as np


class FloatImageRec(C.Structure):          # reference src/V1/klt_util.h:8-12
    _fields_ = [("ncols", C.c_int), ("nrows", C.c_int), ("data", C.POINTER(C.c_float))]


class PyramidRec(C.Structure):             # reference src/V1/pyramid.h:9-14
    _fields_ = [("subsampling", C.c_int), ("nLevels", C.c_int),
                ("img", C.POINTER(C.POINTER(FloatImageRec))),
                ("ncols", C.POINTER(C.c_int)), ("nrows", C.POINTER(C.c_int))]


FI = C.POINTER(FloatImageRec)
PY = C.POINTER(PyramidRec)


class RefLib:
    def __init__(self, capi, path):
        self.capi = capi
        self.api = capi.KLTLibrary(path)
        L = self.lib = self.api.lib
        L._KLTCreateFloatImage.restype = FI
        L._KLTCreateFloatImage.argtypes = [C.c_int, C.c_int]
        L._KLTFreeFloatImage.argtypes = [FI]
        L._KLTToFloatImage.argtypes = [C.c_void_p, C.c_int, C.c_int, FI]
        L._KLTComputeSmoothedImage.argtypes = [FI, C.c_float, FI]
        L._KLTComputeGradients.argtypes = [FI, C.c_float, FI, FI]
        L._KLTGetKernelWidths.argtypes = [C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L._KLTCreatePyramid.restype = PY
        L._KLTCreatePyramid.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L._KLTComputePyramid.argtypes = [FI, PY, C.c_float]
        L._KLTFreePyramid.argtypes = [PY]
        L._quicksort.argtypes = [C.c_void_p, C.c_int]
        self.api.KLTSetVerbosity(0)

    # float image helpers
    def fimg(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        im = self.lib._KLTCreateFloatImage(arr.shape[1], arr.shape[0])
        C.memmove(im.contents.data, arr.ctypes.data, arr.nbytes)
        return im

    def to_np(self, im, free=True):
        w, h = im.contents.ncols, im.contents.nrows
        out = np.ctypeslib.as_array(im.contents.data, shape=(h, w)).copy()
        if free:
            self.lib._KLTFreeFloatImage(im)
        return out

    def kernel_widths(self, sigma):
        a, b = C.c_int(0), C.c_int(0)
        self.lib._KLTGetKernelWidths(sigma, C.byref(a), C.byref(b))
        return a.value, b.value

    def smooth(self, arr, sigma):
        src = self.fimg(arr)
        dst = self.lib._KLTCreateFloatImage(arr.shape[1], arr.shape[0])
        self.lib._KLTComputeSmoothedImage(src, sigma, dst)
        self.lib._KLTFreeFloatImage(src)
        return self.to_np(dst)

    def gradients(self, arr, sigma):
        src = self.fimg(arr)
        gx = self.lib._KLTCreateFloatImage(arr.shape[1], arr.shape[0])
        gy = self.lib._KLTCreateFloatImage(arr.shape[1], arr.shape[0])
        self.lib._KLTComputeGradients(src, sigma, gx, gy)
        self.lib._KLTFreeFloatImage(src)
        return self.to_np(gx), self.to_np(gy)

    def pyramid(self, arr, ss, nlevels, sigma_fact):
        src = self.fimg(arr)
        pyr = self.lib._KLTCreatePyramid(arr.shape[1], arr.shape[0], ss, nlevels)
        self.lib._KLTComputePyramid(src, pyr, sigma_fact)
        out = [self.to_np(pyr.contents.img[l], free=False) for l in range(nlevels)]
        self.lib._KLTFreePyramid(pyr)
        self.lib._KLTFreeFloatImage(src)
        return out

    def last_pyramids(self, tc):
        """(img, gx, gy) level lists of tc->pyramid_last* (sequentialMode)."""
        out = []
        for fld in ("pyramid_last", "pyramid_last_gradx", "pyramid_last_grady"):
            p = C.cast(getattr(tc.contents, fld), PY)
            out.append([self.to_np(p.contents.img[l], free=False) for l in range(p.contents.nLevels)])
        return out

    # public-API conveniences (same shape as the product binding)
    def make_tc(self, **fields):
        tc = self.api.create_context(**fields)
        return tc

    def new_list(self, n):
        return self.api.KLTCreateFeatureList(n)

    def get(self, fl):
        return self.capi.featurelist_to_arrays(fl)

    def put(self, fl, x, y, v):
        self.capi.arrays_to_featurelist(fl, x, y, v)
