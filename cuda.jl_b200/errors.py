"""Exception types mirroring the Julia exceptions the reference path throws."""


class ArgumentError(ValueError):
    """Julia's ArgumentError (accumulate.jl:137,231; sorting.jl:1053)."""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (accumulate.jl:139; Base.check_reducedims, mapreduce.jl:173)."""


class NotGraftedError(TypeError):
    """The call is outside the grafted matrix (user closure, non-neutral init, ...).

    In Julia the new methods `invoke` the reference's generic method in this case
    (SURVEY §8b); there is no reference JIT here, so the mirror fails loudly — it
    never falls back to a CPU or library path.
    """


class B200ArraysError(RuntimeError):
    """Non-zero status from libb200arrays (the analog of CURANDError, lib/curand/src/libcurand.jl:10-32)."""

    def __init__(self, status, text, cuda_error=0):
        self.status = int(status)
        self.cuda_error = int(cuda_error)
        msg = f"{text} (status {status})"
        if cuda_error:
            msg += f" [cudaError {cuda_error}]"
        super().__init__(msg)
