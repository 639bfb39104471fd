"""tf.contrib.slim stand-in: importable (BAISTools imports it); the vgg_16 layers are added where variant B needs them."""
