// empty stand-in (oracle/_ref sensor build only)
#pragma once
