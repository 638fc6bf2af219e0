import numpy as np

def block_reduce(image, block_size=2, func=np.sum, cval=0, func_kwargs=None):
    """Documented skimage behaviour: pad the END of each axis with cval up to a multiple of the
    block size, view as blocks, reduce over the block axes."""
    if np.isscalar(block_size):
        block_size = (block_size,) * image.ndim
    pad = [(0, (-image.shape[i]) % block_size[i]) for i in range(image.ndim)]
    image = np.pad(image, pad, mode="constant", constant_values=cval)
    shp = []
    for i in range(image.ndim):
        shp += [image.shape[i] // block_size[i], block_size[i]]
    blocked = image.reshape(shp)
    return func(blocked, axis=tuple(range(1, 2 * image.ndim, 2)))
