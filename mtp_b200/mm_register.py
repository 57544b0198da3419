"""Import side effect for OpenMMLab configs: ``custom_imports = dict(imports=['mtp_b200.mm_register'])`` registers the
B200-native ``RVSA_MTP`` / ``RVSA_MTP_branches`` in every installed toolkit registry."""
from .registry import register_all

REGISTERED = register_all()
