// Stand-in for the cmake-generated <libgcm/config.hpp> (reference: cmake/config.hpp.in).
// TEST INFRASTRUCTURE ONLY: lets the unmodified reference sources compile in this image.
#pragma once
#define CONFIG_INSTALL_PREFIX ""
#define CONFIG_SHARE_GCM ""
#define CONFIG_ENABLE_LOGGING 0
#define CONFIG_ENABLE_LOGGING_FATAL 0
#define CONFIG_ENABLE_LOGGING_ERROR 0
#define CONFIG_ENABLE_LOGGING_WARN 0
#define CONFIG_ENABLE_LOGGING_INFO 0
#define CONFIG_ENABLE_LOGGING_DEBUG 0
#define CONFIG_ENABLE_LOGGING_TRACE 0
#define CONFIG_ENABLE_ASSERTIONS 1
#ifndef LIBGCM_DOUBLE_PRECISION
#define LIBGCM_DOUBLE_PRECISION 1
#endif
