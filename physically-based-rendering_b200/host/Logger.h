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

class Logger {

	public:
		static int getIndent();
		static int indent( int indent );
		static void logDebug( const char* msg, const char* prefix = "* " );
		static void logDebug( std::string msg, const char* prefix = "* " );
		static void logDebugVerbose( const char* msg, const char* prefix = "* " );
		static void logDebugVerbose( std::string msg, const char* prefix = "* " );
		static void logError( const char* msg, const char* prefix = "* " );
		static void logError( std::string msg, const char* prefix = "* " );
		static void logInfo( const char* msg, const char* prefix = "* " );
		static void logInfo( std::string msg, const char* prefix = "* " );
		static void logWarning( const char* msg, const char* prefix = "* " );
		static void logWarning( std::string msg, const char* prefix = "* " );

	private:
		static void emit( int minLevel, const char* color, const char* msg, const char* prefix );
		static int mIndent;

};

#endif
