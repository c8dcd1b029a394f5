"""Import-path shim: `import upsnet...` resolves to the B200-native implementation (upsnet_b200) under
the reference's own module paths, so model code written against uber-research/UPSNet
(`from upsnet.operators.modules.deform_conv import DeformConv`, `from upsnet.models import *`, ...)
binds to the new engine unchanged."""
