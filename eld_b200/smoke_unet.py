"""U-Net leg of __graft_entry__.smoke(); filled in as the U-Net kernels land."""


def run():
    pass
