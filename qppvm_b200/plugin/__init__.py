"""C++ drop-ins for the two XBotCore RT plugins of the reference, over the C-ABI."""
