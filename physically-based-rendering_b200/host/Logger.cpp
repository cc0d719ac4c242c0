#include "Logger.h"

#include <stdio.h>

int Logger::mIndent = 0;

int Logger::getIndent() {
	return mIndent;
}

int Logger::indent( int indent ) {
	mIndent = ( indent < 0 ) ? 0 : indent;
	return mIndent;
}

void Logger::emit( int minLevel, const char* color, const char* msg, const char* prefix ) {
	int level = 1;
	try {
		level = Cfg::get().value<int>( Cfg::LOG_LEVEL );
	}
	catch( ... ) {}
	if( level < minLevel ) {
		return;
	}
	fprintf( stderr, "%s%*s%s%s\033[0m\n", color, mIndent, "", prefix, msg );
}

void Logger::logDebug( const char* msg, const char* prefix ) { emit( 3, "\033[36m", msg, prefix ); }
void Logger::logDebug( std::string msg, const char* prefix ) { emit( 3, "\033[36m", msg.c_str(), prefix ); }
void Logger::logDebugVerbose( const char* msg, const char* prefix ) { emit( 4, "\033[36m", msg, prefix ); }
void Logger::logDebugVerbose( std::string msg, const char* prefix ) { emit( 4, "\033[36m", msg.c_str(), prefix ); }
void Logger::logError( const char* msg, const char* prefix ) { emit( 1, "\033[31;1m", msg, prefix ); }
void Logger::logError( std::string msg, const char* prefix ) { emit( 1, "\033[31;1m", msg.c_str(), prefix ); }
void Logger::logInfo( const char* msg, const char* prefix ) { emit( 2, "", msg, prefix ); }
void Logger::logInfo( std::string msg, const char* prefix ) { emit( 2, "", msg.c_str(), prefix ); }
void Logger::logWarning( const char* msg, const char* prefix ) { emit( 1, "\033[33m", msg, prefix ); }
void Logger::logWarning( std::string msg, const char* prefix ) { emit( 1, "\033[33m", msg.c_str(), prefix ); }
