"""The batch-file loop of the reference, data path only (controller.processFiles, controller/controller.go:2985-3107): what
sits either side of the block loop is on the device too -
    resample.Time of every input whose rate differs from the target rate      controller.go:2991-3003   -> gdg_resample_time
    common length, rounded up to whole blocks of 8192, zero-padded            controller.go:3005-3031
    block by block: controller.process                                        controller.go:3076-3107   -> gdg_controller_submit_encoded
    Channel.WriteFloats of every output (samplesToBytes)                      controller.go:3118-3143   -> fused into the kernels' store
with two blocks in flight (block k + 1 goes up while block k comes down).  The RIFF container, prompts and file I/O stay
with the caller (out of scope): inputs are float64 arrays (Channel.Floats()), outputs are data-chunk payloads.
"""
import ctypes as C

import numpy as np

from . import FORMATS, GdgError, resample_time

BLOCK_SIZE = 8192                       # controller/controller.go:36
MORE_OUTPUTS_THAN_INPUTS = 3            # controller/controller.go:37: master left / right and the metronome port


def process_files(chains, spat, inputs, rates, target_rate, out_format="PCM16", tuner=None, tuner_channel=-1, meter=None):
    """inputs[c]: float64 samples of channel c at rates[c] Hz.  Returns len(chains) + 3 payloads (bytes) in `out_format`:
    the channels' outputs, master left, master right (zeros without a spatializer) and the metronome port (no metronome on
    this path: zeros), each padded to whole blocks like the reference's output files."""
    lib = chains[0]._lib
    c_in = len(chains)
    if len(inputs) != c_in or len(rates) != c_in:
        raise GdgError("process_files: %d chains, %d inputs, %d rates" % (c_in, len(inputs), len(rates)))
    xs = []
    for x, r in zip(inputs, rates):                                                   # controller.go:2991-3003
        x = np.ascontiguousarray(x, dtype=np.float64)
        xs.append(resample_time(x, int(r), int(target_rate), lib) if int(r) != int(target_rate) else x)
    max_len = max([len(x) for x in xs] + [0])
    if max_len % BLOCK_SIZE:                                                          # controller.go:3016-3018
        max_len = BLOCK_SIZE * (max_len // BLOCK_SIZE + 1)
    fo = FORMATS.index(out_format)
    bo = lib.gdg_format_bytes_per_sample(fo)
    n_rows = c_in + (2 if spat is not None else 0)
    x = np.zeros((c_in, max(max_len, 1)))                                             # zero-padded copies, controller.go:3023-3031
    for c, v in enumerate(xs):
        x[c, :len(v)] = v
    out = np.zeros((n_rows, max(max_len, 1) * bo), dtype=np.uint8)
    handles = (C.c_void_p * c_in)(*[ch._h for ch in chains])
    prev = None
    for k in range(max_len // BLOCK_SIZE):                                            # controller.go:3076-3107
        ip = (C.c_void_p * c_in)(*[x[c].ctypes.data + 8 * BLOCK_SIZE * k for c in range(c_in)])
        op = (C.c_void_p * n_rows)(*[out[r].ctypes.data + bo * BLOCK_SIZE * k for r in range(n_rows)])
        t = lib.gdg_controller_submit_encoded(handles, c_in, tuner._h if tuner else None, tuner_channel, spat._h if spat else None,
                                              meter._h if meter else None, ip, FORMATS.index("IEEE64"), op, fo, n_rows, BLOCK_SIZE, int(target_rate))
        if t < 0:
            raise GdgError(lib.gdg_last_error().decode())
        if prev is not None and lib.gdg_controller_wait(prev) < 0:
            raise GdgError(lib.gdg_last_error().decode())
        prev = t
    if prev is not None and lib.gdg_controller_wait(prev) < 0:
        raise GdgError(lib.gdg_last_error().decode())
    silence = _encode_zeros(lib, fo, max_len)
    rows = [out[r, :max_len * bo].tobytes() for r in range(n_rows)]
    return rows + [silence] * (c_in + MORE_OUTPUTS_THAN_INPUTS - n_rows)


def _encode_zeros(lib, fmt, n):
    """samplesToBytes of n zero samples: all-zero bytes except LPCM8, whose zero is the offset 128 (wave/wave.go:275-311)"""
    return (b"\x80" if FORMATS[fmt] == "PCM8" else b"\x00") * (n * lib.gdg_format_bytes_per_sample(fmt))
