from . import decoder

decoder_dict = {'point': decoder.POINT}          # same registry key as the reference (conv_onet/models/__init__.py:4-6)
