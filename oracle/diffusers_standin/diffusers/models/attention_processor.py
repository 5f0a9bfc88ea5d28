class AttnProcessor2_0:
    pass
