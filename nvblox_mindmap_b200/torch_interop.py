"""Zero-copy glue between torch tensors and raw device pointers."""
import torch

_TYPESTR = {torch.float32: '<f4', torch.float16: '<f2', torch.int32: '<i4', torch.uint8: '|u1'}


class _DeviceBlob:
    """Minimal __cuda_array_interface__ holder; keeps `owner` (the mapper handle) alive."""

    def __init__(self, ptr, shape, dtype, strides_elems, owner):
        item = torch.empty((), dtype=dtype).element_size()
        self._owner = owner
        self.__cuda_array_interface__ = {
            'shape': tuple(int(s) for s in shape),
            'typestr': _TYPESTR[dtype],
            'data': (int(ptr), False),
            'version': 2,
            'strides': None if strides_elems is None else tuple(int(s) * item for s in strides_elems),
        }


def device_view(ptr, shape, dtype, device_index, strides_elems=None, owner=None) -> torch.Tensor:
    """Non-owning tensor view over device memory owned by the mapper handle (from_blob equivalent)."""
    n = 1
    for s in shape:
        n *= int(s)
    if n == 0 or not ptr:
        return torch.empty(tuple(shape), dtype=dtype, device=f'cuda:{device_index}')
    return torch.as_tensor(_DeviceBlob(ptr, shape, dtype, strides_elems, owner), device=f'cuda:{device_index}')


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def current_stream_ptr(device_index: int) -> int:
    """cudaStream_t of torch's current stream on `device_index` (the stream all our kernels ride on)."""
    if _raw_stream is not None:
        return int(_raw_stream(device_index))
    return int(torch.cuda.current_stream(device_index).cuda_stream)
