/*
 * Logger -- reference: source/Logger.{h,cpp}.  Same static interface and level gating
 * (logging.level: 0 none, 1 errors+warnings, 2 +info, 3 +debug, 4 +verbose; config.json:60-67),
 * written to stderr so that a headless driver can keep stdout for data.
 */
#ifndef LOGGER_H
#define LOGGER_H

#include <string>

#include "Cfg.h"

#define LOG_INDENT 4

/* Every level comes as a (const char*) and a (std::string) overload, like upstream. */
#define PBR_LOGGER_LEVEL( name ) \
	static void name( const char* msg, const char* prefix = "* " ); \
	static void name( std::string msg, const char* prefix = "* " );

class Logger {
	public:
		PBR_LOGGER_LEVEL( logError )
		PBR_LOGGER_LEVEL( logWarning )
		PBR_LOGGER_LEVEL( logInfo )
		PBR_LOGGER_LEVEL( logDebug )
		PBR_LOGGER_LEVEL( logDebugVerbose )

		static int indent( int indent );     /* set the indentation of the following lines; returns it */
		static int getIndent();

	private:
		static void emit( int minLevel, const char* color, const char* msg, const char* prefix );
		static int mIndent;
};

#undef PBR_LOGGER_LEVEL

#endif
