"""b200arrays — B200-native reduce / scan / sort for CUDA.jl's CuArray hot path.

Host-side mirror (Python, because no Julia toolchain exists in this image) of the
reference's operator interface for this path: CUDACore/src/mapreduce.jl,
accumulate.jl and sorting.jl.  All device work goes through the C ABI of
libb200arrays.so (include/b200arrays.h); there is no CPU or library fallback.
"""
from . import dtypes
from .errors import ArgumentError, DimensionMismatch, NotGraftedError, B200ArraysError
from .cuarray import CuArray, Array
from .mapreduce import (identity, abs2, add_sum, mul_prod, add, mul, max_, min_, and_, or_, extrema_rf,
                        neutral_element, true_neutral, mapreducedim_, mapreduce, reduce, sum, prod, maximum, minimum,
                        extrema, count, any, all, sum_, prod_, maximum_, minimum_,
                        findmax, findmin, argmax, argmin)
from .accumulate import scan_, accumulate_, accumulate, cumsum, cumprod, cumsum_, cumprod_
from .sorting import (isless, BitonicSort, QuickSort, sort_alg_, sort_, sort, partialsort_alg_, partialsort_, partialsort,
                      sortperm_, sortperm)
from .indexing import findall
from . import multigpu
from . import _lib

__all__ = [n for n in dir() if not n.startswith("_")]
