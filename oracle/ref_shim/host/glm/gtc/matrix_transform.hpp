/* stand-in: nothing of glm/gtc/matrix_transform.hpp is used by the files compiled for the tests */
#include "../glm.hpp"
