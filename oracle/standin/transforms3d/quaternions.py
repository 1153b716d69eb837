def mat2quat(M):  # only imported by the (out of scope) dither module
    raise NotImplementedError('stand-in: quaternions are not on the hot path')
