/* TEST INFRASTRUCTURE: empty stand-in, see ../../../gil.hpp (no PNG encoding in the harness) */
#pragma once
