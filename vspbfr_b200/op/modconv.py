"""placeholder — filled in with the tcgen05 convolution host side."""


def plain_conv_supported(*a, **k):
    return False
